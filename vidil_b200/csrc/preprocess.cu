// Frame pre-processing on the device: uint8 HWC frames -> PIL-identical antialiased bicubic resize (BLIP: to S x S; CLIP:
// shortest edge to S, then the centre S x S window) -> [0,1] -> (x - mean) / std -> fp32 NCHW.
// Reference: run_visual_tokenization.py:138-140 (HF CLIPProcessor, PIL backend) and run_video_CapFilt.py:128-137 (process_frame: torchvision ToPILImage, Resize
// BICUBIC, ToTensor, Normalize), whose resize arithmetic is Pillow's src/libImaging/Resample.c: double-precision weights
// rounded to 22-bit fixed point, int32 accumulation from 1 << 21, >> 22, clip to [0, 255], horizontal pass rounded to uint8
// before the vertical pass.  The integer work is reproduced exactly (bit-identical output); the weight tables are computed
// on the host with the same double-precision operations Pillow performs.
//
// Byte/integer work bound by HBM: pass 1 stages one input row in shared memory (16-byte coalesced loads) and writes one
// resized row; pass 2 reads the few source rows of an output row with consecutive threads on consecutive pixels and writes
// the three fp32 planes coalesced.
#include <math.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "kernels.h"

namespace vidil {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

struct Coeffs {
    int ksize = 0;
    std::vector<int32_t> table;  // [out][2 + ksize]: xmin, count, weights
};

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
const Coeffs& coeffs_for(int in_size, int out_size) {
    static std::map<std::pair<int, int>, Coeffs> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({in_size, out_size});
    if (it != cache.end()) return it->second;
    Coeffs c;
    double scale = static_cast<double>(in_size) / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    c.ksize = static_cast<int>(ceil(support)) * 2 + 1;
    c.table.assign(static_cast<size_t>(out_size) * (2 + c.ksize), 0);
    const double ss = 1.0 / filterscale;
    std::vector<double> w(c.ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = bicubic((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        int32_t* row = c.table.data() + static_cast<size_t>(xx) * (2 + c.ksize);
        row[0] = xmin;
        row[1] = xmax;
        for (int x = 0; x < xmax; ++x) {
            const double v = (ww != 0.0) ? w[x] / ww : w[x];
            row[2 + x] = (v < 0) ? static_cast<int32_t>(-0.5 + v * (1 << PRECISION_BITS))
                                 : static_cast<int32_t>(0.5 + v * (1 << PRECISION_BITS));
        }
    }
    return cache.emplace(std::make_pair(in_size, out_size), std::move(c)).first->second;
}

__device__ __forceinline__ uint8_t clip8(int acc) {
    const int v = acc >> PRECISION_BITS;  // arithmetic shift, as Pillow's lookup index
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One block per (frame, source row): [W,3] -> [S,3], output columns left .. left+S-1 of the resized row (centre crop).
__global__ void __launch_bounds__(256)
    resize_rows_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ tmp, const int32_t* __restrict__ tab, int ksize,
                       int W, int S, int left) {
    extern __shared__ uint8_t srow[];
    const int64_t r = blockIdx.x;
    const uint8_t* src = in + r * W * 3;
    const int nbytes = W * 3;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int i = threadIdx.x; i < nbytes / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(srow)[i] = reinterpret_cast<const uint4*>(src)[i];
        for (int i = (nbytes / 16) * 16 + threadIdx.x; i < nbytes; i += blockDim.x) srow[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < nbytes; i += blockDim.x) srow[i] = src[i];
    }
    __syncthreads();
    uint8_t* dst = tmp + r * S * 3;
    for (int o = threadIdx.x; o < S * 3; o += blockDim.x) {
        const int xx = o / 3, c = o - xx * 3;
        const int32_t* row = tab + static_cast<int64_t>(xx + left) * (2 + ksize);
        const int xmin = row[0], n = row[1];
        int acc = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < n; ++x) acc += static_cast<int>(srow[(xmin + x) * 3 + c]) * row[2 + x];
        dst[o] = clip8(acc);
    }
}

// uint8 -> [0,1]: torchvision's ToTensor divides in float32; transformers' rescale multiplies by the double 1/255 and casts
__device__ __forceinline__ float unit_scale(uint8_t v, bool f64) {
    return f64 ? __double2float_rn(__dmul_rn(static_cast<double>(v), 1.0 / 255.0)) : __fdiv_rn(static_cast<float>(v), 255.0f);
}

// One thread per output pixel (all three channels): vertical pass (output rows top .. top+S_h-1 of the resized image) +
// rescale + normalise.
__global__ void __launch_bounds__(256)
    resize_cols_normalize_kernel(const uint8_t* __restrict__ tmp, float* __restrict__ out, const int32_t* __restrict__ tab,
                                 int ksize, int H, int S_w, int S_h, int top, bool rescale_f64, float m0, float m1, float m2,
                                 float s0, float s1, float s2) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    const int64_t b = blockIdx.z;
    if (xx >= S_w) return;
    const int32_t* row = tab + static_cast<int64_t>(yy + top) * (2 + ksize);
    const int ymin = row[0], n = row[1];
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
    const uint8_t* p = tmp + ((b * H + ymin) * S_w + xx) * 3;
    for (int y = 0; y < n; ++y) {
        const int k = row[2 + y];
        a0 += static_cast<int>(p[0]) * k;
        a1 += static_cast<int>(p[1]) * k;
        a2 += static_cast<int>(p[2]) * k;
        p += static_cast<int64_t>(S_w) * 3;
    }
    // ToTensor: uint8 -> float32 / 255; Normalize: (x - mean) / std, each a single IEEE float32 operation (no FMA contraction)
    const float v0 = __fdiv_rn(__fsub_rn(unit_scale(clip8(a0), rescale_f64), m0), s0);
    const float v1 = __fdiv_rn(__fsub_rn(unit_scale(clip8(a1), rescale_f64), m1), s1);
    const float v2 = __fdiv_rn(__fsub_rn(unit_scale(clip8(a2), rescale_f64), m2), s2);
    const int64_t plane = static_cast<int64_t>(S_h) * S_w;
    float* o = out + b * 3 * plane + static_cast<int64_t>(yy) * S_w + xx;
    o[0] = v0;
    o[plane] = v1;
    o[2 * plane] = v2;
}

inline size_t align1k(size_t x) { return (x + 1023) / 1024 * 1024; }

// Geometry of one pre-processing recipe: resize the [H, W] frame to [rh, rw], keep the S x S window at (top, left).
struct Geometry {
    int rh, rw, top, left;
    bool rescale_f64;
};
// run_video_CapFilt.py:128-137: Resize((S, S)) — both sides to S, no crop, ToTensor's float32 division.
Geometry blip_geometry(int, int, int S) { return Geometry{S, S, 0, 0, false}; }
// transformers' CLIPImageProcessor (PIL backend) as run_visual_tokenization.py:138-140 calls it: shortest edge to S keeping
// the aspect ratio (the long side is int(S * long / short), image_transforms.get_resize_output_image_size), centre crop
// S x S at ((rh - S) // 2, (rw - S) // 2), rescale by the double 1/255.
Geometry clip_geometry(int H, int W, int S) {
    Geometry g;
    if (W <= H) {
        g.rw = S;
        g.rh = static_cast<int>(static_cast<int64_t>(S) * H / W);
    } else {
        g.rh = S;
        g.rw = static_cast<int>(static_cast<int64_t>(S) * W / H);
    }
    g.top = (g.rh - S) / 2;
    g.left = (g.rw - S) / 2;
    g.rescale_f64 = true;
    return g;
}

size_t workspace_for(int B, int H, int W, int S, const Geometry& g) {
    const Coeffs& ch = coeffs_for(W, g.rw);
    const Coeffs& cv = coeffs_for(H, g.rh);
    return align1k(static_cast<size_t>(B) * H * S * 3) + align1k(ch.table.size() * 4) + align1k(cv.table.size() * 4);
}

int run_geometry(const uint8_t* frames, int B, int H, int W, int S, const Geometry& g, const float* mean, const float* stdv,
                 float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (B <= 0) return 0;
    if (H <= 0 || W <= 0 || S <= 0 || W * 3 > 200 * 1024) {
        set_error("preprocess: unsupported geometry H=%d W=%d S=%d", H, W, S);
        return 1;
    }
    if (workspace_bytes < workspace_for(B, H, W, S, g) || (reinterpret_cast<uintptr_t>(workspace) & 1023)) {
        set_error("preprocess: workspace too small or not 1024-byte aligned");
        return 1;
    }
    const Coeffs& ch = coeffs_for(W, g.rw);
    const Coeffs& cv = coeffs_for(H, g.rh);
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    uint8_t* tmp = base;
    int32_t* tab_h = reinterpret_cast<int32_t*>(base + align1k(static_cast<size_t>(B) * H * S * 3));
    int32_t* tab_v = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(tab_h) + align1k(ch.table.size() * 4));
    VIDIL_CUDA_OK(cudaMemcpyAsync(tab_h, ch.table.data(), ch.table.size() * 4, cudaMemcpyHostToDevice, stream));
    VIDIL_CUDA_OK(cudaMemcpyAsync(tab_v, cv.table.data(), cv.table.size() * 4, cudaMemcpyHostToDevice, stream));
    const size_t smem = static_cast<size_t>(W) * 3 + 16;
    // per launch (cheap): the attribute belongs to the current device's copy of the function
    VIDIL_CUDA_OK(cudaFuncSetAttribute(resize_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16));
    resize_rows_kernel<<<B * H, 256, smem, stream>>>(frames, tmp, tab_h, ch.ksize, W, S, g.left);
    VIDIL_CUDA_OK(cudaGetLastError());
    const dim3 grid((S + 255) / 256, S, B);
    resize_cols_normalize_kernel<<<grid, 256, 0, stream>>>(tmp, out, tab_v, cv.ksize, H, S, S, g.top, g.rescale_f64, mean[0],
                                                          mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(2);
    return 0;
}

}  // namespace

size_t preprocess_workspace_bytes(int B, int H, int W, int S) {
    if (B <= 0 || H <= 0 || W <= 0 || S <= 0) return 0;
    return workspace_for(B, H, W, S, blip_geometry(H, W, S));
}
size_t clip_preprocess_workspace_bytes(int B, int H, int W, int S) {
    if (B <= 0 || H <= 0 || W <= 0 || S <= 0) return 0;
    return workspace_for(B, H, W, S, clip_geometry(H, W, S));
}

int preprocess_run(const uint8_t* frames, int B, int H, int W, int S, const float* mean, const float* stdv, float* out,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    return run_geometry(frames, B, H, W, S, blip_geometry(H, W, S), mean, stdv, out, workspace, workspace_bytes, stream);
}
int clip_preprocess_run(const uint8_t* frames, int B, int H, int W, int S, const float* mean, const float* stdv, float* out,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    return run_geometry(frames, B, H, W, S, clip_geometry(H, W, S), mean, stdv, out, workspace, workspace_bytes, stream);
}

}  // namespace vidil
