/* ORACLE (test infrastructure, never on the product path).
 *
 * Plain-C restatement of the reference's RayDirectionComputer, whose own source needs Eigen3
 * (src/preprocessing/ray_direction_computer.h:4-5), which is not installed: "unbuildable here".
 * The arithmetic is scalar float and is restated line for line.
 * PARITY: unpinned by the reference (it has no test or golden vector for this class; the
 * test_ray_directions target named in README.md:289 does not exist).  Pinned instead by
 * known-answer cases in tests/test_rays.py.
 *
 * Build: gcc -O2 -ffp-contract=off.  The reference is built -O3 -march=native (CMakeLists.txt:15-17),
 * where GCC may or may not contract x*x+y*y+z*z into FMAs depending on the host; the un-contracted
 * form is the one the source states, and the CUDA kernel is compared to it at <= 2 ulp.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

/* computeRayDirections: ray_direction_computer.cpp:17-62.  K row-major 3x3.  out: (H*W,3). */
void oracle_rays_hw3(const float* K, int height, int width, float* out) {
    float fx = K[0], fy = K[4], cx = K[2], cy = K[5];     /* :29-32 */
    float fx_inv = 1.0f / fx;                              /* :35 */
    float fy_inv = 1.0f / fy;                              /* :36 */
    int idx = 0;
    for (int v = 0; v < height; ++v) {                     /* :40 */
        for (int u = 0; u < width; ++u) {                  /* :41 */
            float x = ((float)u - cx) * fx_inv;            /* :47 */
            float y = ((float)v - cy) * fy_inv;            /* :48 */
            float z = 1.0f;                                /* :49 */
            float norm = sqrtf(x * x + y * y + z * z);     /* :52 */
            out[3 * idx + 0] = x / norm;                   /* :53 */
            out[3 * idx + 1] = y / norm;                   /* :54 */
            out[3 * idx + 2] = z / norm;                   /* :55 */
            ++idx;
        }
    }
}

/* computeRayDirectionsMaps: ray_direction_computer.cpp:64-101.  out: (3,H,W) planar. */
void oracle_rays_3hw(const float* K, int height, int width, float* out) {
    float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    float fx_inv = 1.0f / fx, fy_inv = 1.0f / fy;
    size_t hw = (size_t)height * width;
    for (int v = 0; v < height; ++v)
        for (int u = 0; u < width; ++u) {
            float x = ((float)u - cx) * fx_inv;
            float y = ((float)v - cy) * fy_inv;
            float z = 1.0f;
            float norm = sqrtf(x * x + y * y + z * z);
            size_t i = (size_t)v * width + u;
            out[i] = x / norm;
            out[hw + i] = y / norm;
            out[2 * hw + i] = z / norm;
        }
}

/* transformRaysToWorld: ray_direction_computer.cpp:103-127.  pose row-major 4x4; rays (N,3). */
void oracle_rays_to_world(const float* rays, int64_t n, const float* pose, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        float a = rays[3 * i], b = rays[3 * i + 1], c = rays[3 * i + 2];
        float w0 = pose[0] * a + pose[1] * b + pose[2] * c;    /* R * r, :115 */
        float w1 = pose[4] * a + pose[5] * b + pose[6] * c;
        float w2 = pose[8] * a + pose[9] * b + pose[10] * c;
        float nrm = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);        /* Eigen normalize(), :119 */
        if (nrm > 0.0f) { w0 /= nrm; w1 /= nrm; w2 /= nrm; }
        out[3 * i] = w0; out[3 * i + 1] = w1; out[3 * i + 2] = w2;
    }
}

/* saveRayDirections: ray_direction_computer.cpp:129-168.  int32 H, int32 W, H*W*3 float32. */
int oracle_rays_save(const char* filename, const float* rays, int height, int width) {
    FILE* f = fopen(filename, "wb");
    if (!f) return 0;
    int32_t h = height, w = width;
    fwrite(&h, sizeof(int32_t), 1, f);
    fwrite(&w, sizeof(int32_t), 1, f);
    fwrite(rays, sizeof(float), (size_t)height * width * 3, f);
    fclose(f);
    return 1;
}

/* loadRayDirections: ray_direction_computer.cpp:170-201.  Returns malloc'd (H*W,3) or NULL. */
float* oracle_rays_load(const char* filename, int* height, int* width) {
    FILE* f = fopen(filename, "rb");
    if (!f) return NULL;
    int32_t h = 0, w = 0;
    if (fread(&h, sizeof(int32_t), 1, f) != 1 || fread(&w, sizeof(int32_t), 1, f) != 1) { fclose(f); return NULL; }
    size_t n = (size_t)h * w * 3;
    float* r = (float*)malloc(n * sizeof(float));
    if (fread(r, sizeof(float), n, f) != n) { free(r); fclose(f); return NULL; }
    fclose(f);
    *height = h; *width = w;
    return r;
}

void oracle_free(void* p) { free(p); }
