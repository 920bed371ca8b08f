// k_preprocess.cuh -- the verifier's val/test pre-processing, fused: resize (cv2 INTER_LINEAR fixed point) -> centre crop ->
// HWC u8 -> CHW f32 -> normalise -> channel concatenation of the four renders of a hypothesis.
//
// Reference chain (salve/train_utils.py:126-159):
//   ResizeQuadruplet     salve/utils/transform.py:256-272   cv2.resize(img, (w, h), INTER_LINEAR)
//   CropQuadruplet       transform.py:386-420               centre, offset int((resize - crop) / 2)
//   ToTensorQuadruplet   transform.py:79-85, 105-123        HWC -> CHW float32 (no /255)
//   NormalizeQuadruplet  transform.py:177-202               t.sub_(m).div_(s), ImageNet mean/std * 255 (normalization_utils.py:13-26)
//   torch.cat([x1, x2, x3, x4], dim=1)                      salve/models/early_fusion.py:60-61, order zind_data.py:306-315
// cv2's 8-bit bilinear is fixed point (imgproc/src/resize.cpp): 11-bit weights, horizontal pass in int, vertical pass
//   (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2
// reproduced exactly here, so the result is bit-identical to the reference chain (weights come from the host, computed with
// cv2's own float recipe).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bev {

struct ResizeTap { int32_t s0, s1, w0, w1; };  // source indices and 11-bit weights of one destination index

struct PreprocArgs {
    const uint8_t* const* src;  // n * 4 device pointers: x1c, x2c, x1f, x2f of each hypothesis (src_h x src_w x 3 u8)
    const ResizeTap* xtap;      // crop_w entries (crop offset applied)
    const ResizeTap* ytap;      // crop_h entries
    int32_t src_w, crop_h, crop_w;
    float mean[3], stdv[3];
    float* out;                 // n x 12 x crop_h x crop_w
};

constexpr int PREPROC_ROWS = 8;  // output rows per CTA

// grid = (ceil(crop_h / PREPROC_ROWS), 4, n), block = crop_w rounded up to a warp (<= 1024)
__global__ void __launch_bounds__(256) verifier_preprocess_kernel(PreprocArgs A) {
    const int x = threadIdx.x;
    const int k = blockIdx.y, n = blockIdx.z;
    const uint8_t* img = A.src[n * 4 + k];
    const int y0 = blockIdx.x * PREPROC_ROWS;
    if (x >= A.crop_w) return;
    const ResizeTap tx = A.xtap[x];
    const int o0 = tx.s0 * 3, o1 = tx.s1 * 3;
    const size_t plane = (size_t)A.crop_h * A.crop_w;
    float* outp = A.out + ((size_t)n * 12 + (size_t)k * 3) * plane + x;
    const size_t pitch = (size_t)A.src_w * 3;
#pragma unroll 2
    for (int yy = 0; yy < PREPROC_ROWS; yy++) {
        const int y = y0 + yy;
        if (y >= A.crop_h) break;
        const ResizeTap ty = A.ytap[y];
        const uint8_t* r0 = img + (size_t)ty.s0 * pitch;
        const uint8_t* r1 = img + (size_t)ty.s1 * pitch;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int h0 = (int)r0[o0 + c] * tx.w0 + (int)r0[o1 + c] * tx.w1;
            const int h1 = (int)r1[o0 + c] * tx.w0 + (int)r1[o1 + c] * tx.w1;
            const int v = (((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
            outp[(size_t)c * plane + (size_t)y * A.crop_w] = __fdiv_rn(__fsub_rn((float)v, A.mean[c]), A.stdv[c]);
        }
    }
}

}  // namespace bev
