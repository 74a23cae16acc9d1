// ob_ir.h -- intermediate representation handed from the symbol-decode kernel (one thread per
// (stream, frame)) to the warp-parallel band-reconstruction kernel and the block-parallel synthesis
// kernel.  Everything in here is a pure integer function of (packet bytes, TOC): see SURVEY.md 3.1,
// "the symbol-decode stage reads no inter-frame state".
//
// Plain C++ (no CUDA types) so the host-side emulation harness in tests/ can include it too.
#pragma once
#include <stdint.h>

#define OB_NB 21              // bands of mode48000_960_120 (opus/celt/static_modes_float.h:866-888)
#define OB_OVERLAP 120
#define OB_SHORT 120          // shortMdctSize
#define OB_MAX_N 960          // samples / channel / frame at 20 ms
#define OB_MAX_BAND 176       // widest band (22 bins * 8)
#define OB_MAX_LEAVES 672     // 21 bands * 2 channels * 2^(LM+1) partitions: cannot be exceeded by construction
#define OB_NORM_LEN 624       // M * eBands[20] at LM=3: folding source per channel (bands.c:1438)

// OPUS_* error codes (src/error.rs:36-62 <-> opus_defines.h)
#define OB_OK 0
#define OB_BAD_ARG (-1)
#define OB_BUFFER_TOO_SMALL (-2)
#define OB_INTERNAL_ERROR (-3)
#define OB_INVALID_PACKET (-4)
#define OB_UNIMPLEMENTED (-5)
#define OB_INVALID_STATE (-6)
#define OB_ALLOC_FAIL (-7)

enum ObLeafKind { OB_LEAF_ZERO = 0, OB_LEAF_PULSES = 1, OB_LEAF_NOISE = 2, OB_LEAF_FOLD = 3, OB_LEAF_ONE = 4 };

// One terminal partition of quant_partition (opus/celt/bands.c:1038-1103) or an N==1 band (:904-937).
struct ObLeaf {
    uint16_t off;         // first coefficient, absolute index into the frame's X (channel c starts at c*N)
    uint8_t n;            // coefficients in this partition (1..176)
    uint8_t K;            // pulses (PULSES) / sign bit (ONE)
    uint8_t kind;         // ObLeafKind
    uint8_t B;            // short blocks spanned (stride of exp_rotation, vq.c:74-117)
    uint16_t lcg_before;  // NOISE / FOLD: celt_lcg_rand steps taken in this frame before this leaf (bands.c:1075,1085); PULSES: the pulse vector's squared norm (<= 255^2)
    float gain;           // product of the mid/side gains down the split tree (bands.c:1023-1034)
};

enum ObBandMode { OB_BAND_MONO = 0, OB_BAND_DUAL = 1, OB_BAND_JOINT = 2, OB_BAND_JOINT_N2 = 3 };

// One band of quant_all_bands (opus/celt/bands.c:1455-1668).  "call a" is quant_band on X (or on x2 for the
// N==2 stereo case), "call b" is quant_band on Y (dual stereo: second channel; joint: the side).
struct ObBand {
    uint16_t leaf_begin_a, leaf_begin_b;
    uint8_t leaf_cnt_a, leaf_cnt_b;
    int16_t eff_lowband;  // folding source offset into norm[] (bands.c:1527) or -1
    int16_t imid, iside;  // bitexact_cos outputs of the band-level stereo split (bands.c:885-897)
    uint8_t mode;         // ObBandMode
    uint8_t flags;        // bit0 inv (bands.c:1374), bit1 N==2: x2 is Y (c=itheta>8192), bit2 N==2: sign bit
    uint8_t pad[2];
};

#define OB_F_SILENCE 1
#define OB_F_TRANSIENT 2
#define OB_F_INTRA 4
#define OB_F_POSTFILTER 8
#define OB_F_ANTICOLLAPSE 16
#define OB_F_LOST 32             // lost packet (len == 0) or DTX payload (<= 1 byte): conceal `status` samples (opus_decoder.c:284-334, :715-729)
#define OB_F_DTX 64              // with OB_F_LOST: a DTX frame of a packet that did arrive (its TOC still sets the decoder's frame size)

struct ObFrameHdr {
    int32_t status;            // samples per channel (>0) or OPUS_* error (<0)
    uint32_t final_range;      // dec->rng at the end of the frame (celt_decoder.c:1358) -> OPUS_GET_FINAL_RANGE
    uint16_t n_leaves;
    uint16_t pf_pitch;         // post-filter period (celt_decoder.c:1145)
    uint8_t LM, C, end, flags;
    uint8_t spread, pf_tapset, pf_qg, coded_bands;
    uint8_t intensity, dual_stereo;
    uint8_t skip_in, end_in;   // written by the plan pass: st->skip_plc and st->end (0 = no packet decoded yet) before this frame
    uint32_t lcg_total;        // LCG steps taken by all bands (seed for anti_collapse = jump(seed_in, lcg_total))
    uint32_t seed_in;          // plan pass: st->rng before this frame (noise fill / folding seed, celt_decoder.c:1279)
    int32_t loss_in;           // plan pass: st->loss_duration before this frame (2.5 ms units, celt_decoder.c:965)
    uint16_t lastfs_in;        // plan pass: OpusDecoder.frame_size before this frame: a concealment never exceeds it (opus_decoder.c:288-289)
    uint16_t pad2;
    int16_t coarse_qi[2 * OB_NB];   // Laplace-decoded coarse energy deltas [c*21+i] (quant_bands.c:450-479)
    int16_t pulses[OB_NB];          // PVQ bit allocation per band, 1/8 bit (anti_collapse depth, bands.c:289)
    uint8_t fine_quant[OB_NB];      // fine energy bits per band (rate.c ebits)
    uint8_t fine_q2[2 * OB_NB];     // fine energy raw values [c*21+i] (quant_bands.c:505)
    int8_t final_bit[2 * OB_NB];    // energy finalise bit [c*21+i]: -1 none, else 0/1 (quant_bands.c:531)
    uint8_t collapse_masks[2 * OB_NB];  // [i*C+c] as in the reference (bands.c:1660-1661)
    int8_t tf_change[OB_NB];        // tf_res after tf_select_table (celt_decoder.c:493-496)
    uint8_t pad1[3];
};

// One CELT frame to decode or conceal, written by the framing pass (opus_packet_parse_impl, opus/src/opus.c:194-353, and the
// frame loop of opus_decode_native, opus_decoder.c:715-799).  A code-0 packet is one slot; a code-1/2/3 packet is `count` slots.
#define OB_SLOT_FIRST 1          // first frame of its packet
#define OB_SLOT_LAST 2           // last frame of its packet: the packet's sample count / final range are written after it
struct ObSlot {
    uint32_t off;                // byte offset of the frame's payload in the packet buffer
    int16_t len;                 // payload bytes (0..1275)
    uint8_t toc;                 // the packet's TOC byte
    uint8_t flags;               // OB_SLOT_*
    uint16_t pkt;                // which of the stream's packets of this call the frame belongs to
    uint16_t sample_off;         // first output sample (per channel) inside that packet's PCM slot
    int32_t status;              // 0: decode `len` bytes; > 0: conceal that many samples (lost packet / DTX frame); < 0: OPUS_* error of the packet
};

// Per-frame IR slot: header, band records, leaves, pulse vector.
struct ObFrameIR {
    ObFrameHdr hdr;
    ObBand bands[OB_NB];
    ObLeaf leaves[OB_MAX_LEAVES];
    int16_t iy[2 * OB_MAX_N];       // signed pulse counts, same indexing as X
};
