"""The JSON line bench.py prints is a contract with the driver (task prompt, 'bench.py keeps the contract below' + tier section 4): these tests
check the committed line of the final tree (profiles/r02final_bench.json, written by `python bench.py` on a B200) and bench.py's CPU-side helpers
-- no GPU needed."""
import json
import os

import pytest

from conftest import ROOT

LINE = os.path.join(ROOT, "profiles", "r02final_bench.json")


def _line():
    return json.loads(open(LINE).read().splitlines()[-1])


def test_committed_bench_line_has_every_contract_key():
    d = _line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] in json.dumps(base) or d["metric"] == "decode_audio_seconds_per_second"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                                          # an end-to-end number of its own, not the device-timed one repeated
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_bench_line_is_self_consistent():
    d = _line()
    S, F = d["config"]["streams_per_gpu"], d["config"]["frames_per_stream_per_step"]
    audio = S * F * d["config"]["frame_ms"] / 1e3 * d["n_gpus"]
    assert abs(d["value"] - audio / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6           # value = the units all ranks processed / the step time
    kms = d["roofline"]["kernel_ms"]
    assert sum(kms.values()) <= d["ms_per_step"] * 1.001                                       # the three kernels fit inside the step
    assert d["roofline"]["kernel"] == max(kms, key=kms.get)                                    # the roofline is quoted for the dominant kernel
    import bench
    assert abs(d["roofline"]["algorithmic_bytes_per_frame"] - bench.algorithmic_bytes_per_frame(F)) < 1e-6
    assert d["e2e"]["d2h_bytes_per_step"] >= S * F * 960 * 4                                    # the float PCM crosses PCIe every step
    assert d["e2e"]["value"] < d["value"] * 1.001
    # traffic: per-frame bytes of the committed ncu capture times this launch's frames
    tj = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))["decode"]
    assert abs(d["roofline"]["traffic"] - tj["kernels"][d["roofline"]["kernel"]]["dram_bytes_per_frame"] * S * F) / d["roofline"]["traffic"] < 1e-6
    for leg in ("decode_mono_64k", "decode_stereo_96k", "encode_stereo_96k_c10"):
        sweep = d["live"][leg]["sweep"]
        rt = [r["streams_per_gpu"] for r in sweep if r["realtime"]]
        assert d["live"][leg]["rt_stream_capacity_per_gpu"] == (max(rt) if rt else 0)
        assert all(r["realtime"] == (r["ms_per_call"] < 20.0) for r in sweep)


def test_algorithmic_bytes_follow_survey_8d():
    import bench
    # B_dec = P + 4*C*N + S_dec/F (config 2: P = 160, C = 1, N = 960, S_dec = 10 368)
    assert bench.algorithmic_bytes_per_frame(50) == pytest.approx(160 + 3840 + 10368 / 50)
    assert bench.algorithmic_bytes_per_frame(1) == pytest.approx(14368)
    assert bench.enc_algorithmic_bytes_per_frame(10) == pytest.approx(7680 + 240 + 19392 / 10)


def test_traffic_file_is_what_the_committed_ncu_summaries_say():
    """profiles/r02_dram_traffic.json (bench.py's roofline.traffic) is generated, not typed: tools/make_dram_traffic.py over the committed ncu summaries
    must reproduce it."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_dram_traffic.py"), "profiles/r02fin_ncu_decoder_kernels.txt",
                          "profiles/r02fin_ncu_encoder_thread_kernel.txt", "profiles/r02fin_ncu_encoder_warp_kernel.txt"], cwd=ROOT, capture_output=True, text=True, check=True).stdout
    assert json.loads(out) == json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))


def test_both_arms_print_the_same_config_object():
    """The driver compares the `config` of `--impl ours` with that of `--impl reference`: they come from one function, and the committed lines of the two
    arms (same box, same tree) carry equal objects."""
    import bench
    c = bench.workload_config(200)
    assert c["workload"] == bench.WORKLOAD and c["streams_per_gpu"] == bench.STREAMS_PER_GPU and c["frames_per_stream_per_step"] == 200
    assert "L2" in c["l2"] and "collective" in c["sharding"]
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r02final_bench_reference.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["config"] == _line()["config"]
    assert ref["e2e"]["value"] == ref["value"] and ref["cpu_baseline"]["kind"] == "reference" and ref["gpu_launches"] == 0
