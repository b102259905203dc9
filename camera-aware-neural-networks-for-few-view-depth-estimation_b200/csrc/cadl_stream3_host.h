// Host-side entry of the round-2 streaming gradient kernel (cadl_stream3.cu), called by cadl_api.cu.
#pragma once
#include <cuda_runtime.h>
#include "cadl_args.cuh"

namespace cadl {

// One CTA of twelve warps per SM.  The same twelve warps as three CTAs of four (the register file holds no more at 166
// registers), but warps of ONE CTA have the same age for the schedulers and progress evenly: with three CTAs the warps of
// the first-dispatched one finish their rows at 40 us and those of the last at 47 (profiles/r02_trace.txt), and the SM's
// last third runs under-occupied.  65.3 -> 63.8 us (4 x 96 threads and 2 x 192: 64.3).
#ifndef CADL_S3_THREADS
#define CADL_S3_THREADS 384
#endif
#ifndef CADL_S3_MINB
#define CADL_S3_MINB 1
#endif
#ifndef CADL_S3_DEPTH
#define CADL_S3_DEPTH 5
#endif
constexpr int kS3Threads = CADL_S3_THREADS;
constexpr int kS3MinBlocks = CADL_S3_MINB;
constexpr int kS3Depth = CADL_S3_DEPTH;          // ring slots per warp: the current row, the next one, Depth - 2 in flight
constexpr int kS3BoxW = 136;                     // pixels per TMA box: 4 halo + 128 + 4 halo
constexpr int kS3BoxBytes = kS3BoxW * 4;
constexpr int kS3C1BoxW = 72;                    // cells per C1 box: 4 + 64 + 4
constexpr int kS3C1BoxBytes = kS3C1BoxW * 4;
constexpr int kS3C1Floats = 96;                  // slot pitch of the C1 row: 384 B
constexpr int kS3RowFloats = 160;                // slot pitch of one tensor's row: 640 B, a multiple of TMA's 128-byte alignment
constexpr int kS3RgbDepth = 3;                   // ONE box (136 x 1 x 3) for the three channel rows: lands dense
constexpr int kS3RgbPitch = kS3BoxW;
constexpr int kS3RgbFloats = (3 * kS3RgbPitch + 31) / 32 * 32;   // 416 (dense) or 480

struct alignas(16) ImgRec {        // per image, zero between calls
    unsigned long long hi[5];      // fixed point, 2^-16 units:  GX0, GY0, SMX, SMY, RP_E of this image
    unsigned long long lo[5];      // fixed point, 2^-56 units (the remainders)
    unsigned int cnt;              // shares finished
    unsigned int flags;            // bit q: quantity q received a NaN partial; bit 8 + q: an infinite one
    float off;                     // a_b * L_b / (HW) * w * upstream      (valid once ready != 0)
    unsigned int ready;
    double Lb;                     // this image's share of the smoothness loss
    unsigned int pad[4];
};
static_assert(sizeof(ImgRec) == 128, "ImgRec layout");
enum { IQ_GX0 = 0, IQ_GY0, IQ_SMX, IQ_SMY, IQ_RP, IQ_COUNT };

struct Stream3Args {
    const float* c1;          // coarse-scale field (B, H/2, W/2) of pyr_coef_kernel
    const unsigned long long* pyr_rec;   // pyr_coef_kernel's fixed-point totals (the record after the B image records:
                                         // hi[6], lo[6], flags): loss sums of scales 1..3
    ImgRec* img;
    unsigned int* done;       // warps finished (returned to 0)
    unsigned int* epoch;      // calls completed on this workspace: the value of ImgRec::ready that means "this call"
    int nstrip;               // 128-column strips per image row
    int kpi;                  // row ranges per strip: a share is (image, strip, row range)
    int nwarps;               // warps that take part
    float lsnx, lsny;         // log2(w_smooth * upstream / N_x), log2(w_smooth * upstream / N_y)
    double inv_snx, inv_sny;  // their inverses: back from the scaled sums to sum w|dp|
    float inx0, iny0;         // w_grad * upstream / (4 N_x), / (4 N_y) of scale 0           depth_loss.h:162-163
    double rnx[4], rny[4];    // 1 / (global_B * Hs * (Ws - 1)), 1 / (global_B * (Hs - 1) * Ws) per scale; inf where a scale has no edges
    double r_hw;              // 1 / (H * W)
};

// F = FB_* mask (15, 7 or FB_GRAD).  Fills the launch geometry of sa, encodes the tensor maps and launches
// cooperatively; returns cudaErrorNotSupported when the device / shape cannot hold the grid co-resident or the
// tensor maps cannot be encoded (the caller then takes another path).
cudaError_t launch_stream3(int F, const PhaseBArgs& a, Stream3Args& sa, cudaStream_t st);
bool stream3_fill_smooth(const PhaseBArgs& a, Stream3Args& sa);
}  // namespace cadl
