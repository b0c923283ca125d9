// emb.cu -- the MAG-Scholar-C side of the hot path beyond the plain aggregation:
//   * MLP.emb with a non-zero input dropout (/root/reference/model_mag.py:48-55): element-wise dropout on the gathered
//     embedding rows, fused into the weighted segment reduction -- forward and backward regenerate the same
//     counter-based Philox mask, no [nza, H] temporary and no stored mask;
//   * the optimizer step of the embedding table (/root/reference/model_mag.py:27,312-313,367-369): the reference keeps
//     a DENSE gradient [2.78 M, H] and lets torch.optim.Adam walk the whole table every step.  Here the gradient stays
//     on the rows the batch touched and Adam is evaluated lazily but EXACTLY: a row that is not touched for k steps
//     still moves under dense Adam (its moments decay, the update m_hat / (sqrt(v_hat) + eps) is not zero), so a row
//     is caught up -- the k zero-gradient steps replayed in registers -- whenever it is read or updated again.
#include "gp_common.cuh"

#include <algorithm>

namespace {

constexpr int kEmbBlock = 256;

// keep decision of element (entry j, column c): Philox block (j, c / 4), lane c % 4
__device__ __forceinline__ void emb_keep4(uint64_t j, uint32_t cblk, uint64_t seed, uint64_t offset, uint32_t thresh, bool keep[4]) {
    uint32_t r[4];
    gp_philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), cblk, (uint32_t)offset, (uint32_t)seed,
                     (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32) ^ 0x9E3779B9u, r);
#pragma unroll
    for (int i = 0; i < 4; i++) keep[i] = r[i] >= thresh;
}

// One warp per output row; lanes own column blocks of 4 (float4 loads), entries are walked one by one.
// out[b,:] = sum_j w_j * drop(E[idx_j,:]) / (sum_j w_j + eps),   drop(x) = keep ? x / (1-p) : 0   (model_mag.py:49-54)
__global__ void emb_dropout_fwd_kernel(const float *table, long long ld_table, int H, const int *row_ptr, const int *idx,
                                       const float *w, long long B, float scale, uint32_t thresh, uint64_t seed,
                                       uint64_t offset, float eps, float *out, long long ld_out, float *denom_out) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int lo = row_ptr[b], hi = row_ptr[b + 1];
    float wsum = 0.f;
    for (int j = lo + lane; j < hi; j += 32) wsum += w[j];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    const float denom = wsum + eps;
    if (lane == 0 && denom_out) denom_out[b] = denom;
    for (int c0 = lane * 4; c0 < H; c0 += 128) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = lo; j < hi; j++) {
            const float wj = __ldg(w + j) * scale;
            const float *row = table + (long long)__ldg(idx + j) * ld_table;
            bool keep[4];
            emb_keep4((uint64_t)j, (uint32_t)(c0 >> 2), seed, offset, thresh, keep);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (c0 + i < H && keep[i]) acc[i] = fmaf(wj, __ldg(row + c0 + i), acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (c0 + i < H) out[b * ld_out + c0 + i] = acc[i] / denom;
    }
}

// grad_rows[slot_j, c] += keep(j,c) * w_j / (1-p) / denom_b * grad_out[b, c];  slot_j = compact row of entry j
__global__ void emb_dropout_bwd_kernel(const float *grad_out, long long ld_go, int H, const int *row_ptr, const int *slot,
                                       const float *w, const float *denom, long long B, float scale, uint32_t thresh,
                                       uint64_t seed, uint64_t offset, float *grad_rows, long long ld_gr) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int lo = row_ptr[b], hi = row_ptr[b + 1];
    const float inv = scale / denom[b];
    for (int c0 = lane * 4; c0 < H; c0 += 128) {
        float g[4];
#pragma unroll
        for (int i = 0; i < 4; i++) g[i] = c0 + i < H ? grad_out[b * ld_go + c0 + i] * inv : 0.f;
        for (int j = lo; j < hi; j++) {
            const float wj = __ldg(w + j);
            float *row = grad_rows + (long long)__ldg(slot + j) * ld_gr;
            bool keep[4];
            emb_keep4((uint64_t)j, (uint32_t)(c0 >> 2), seed, offset, thresh, keep);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (c0 + i < H && keep[i]) atomicAdd(row + c0 + i, wj * g[i]);
        }
    }
}

__global__ void emb_mask_kernel(long long nza, int H, uint32_t thresh, uint64_t seed, uint64_t offset, uint8_t *mask) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nblk = (long long)nza * ((H + 3) / 4);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nblk; t += stride) {
        const long long j = t / ((H + 3) / 4);
        const int cb = (int)(t % ((H + 3) / 4));
        bool keep[4];
        emb_keep4((uint64_t)j, (uint32_t)cb, seed, offset, thresh, keep);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (cb * 4 + i < H) mask[j * H + cb * 4 + i] = keep[i] ? 1 : 0;
    }
}

// ---- exact lazy Adam ------------------------------------------------------------------------------------------------
// torch.optim.Adam (amsgrad off, maximize off, weight_decay 0), one step with gradient g at step number s:
//   m = m + (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g
//   p -= (lr / (1 - b1^s)) * m / (sqrt(v) / sqrt(1 - b2^s) + eps)
struct AdamHyper { float lr, b1, b2, eps; };

__device__ __forceinline__ void adam_one(float &p, float &m, float &v, float g, int s, const AdamHyper &h) {
    m = m + (g - m) * (1.f - h.b1);
    v = v * h.b2 + (1.f - h.b2) * g * g;
    const float bc1 = 1.f - powf(h.b1, (float)s);
    const float bc2 = 1.f - powf(h.b2, (float)s);
    const float step_size = h.lr / bc1;
    const float den = sqrtf(v) / sqrtf(bc2) + h.eps;
    p = p - step_size * (m / den);
}

// Brings the rows rows[0..R) (or every row when rows == NULL) up to date with step `upto` by replaying the zero-gradient
// steps last[row]+1 .. upto, and -- when grad != NULL -- applies step upto+1 with the row's gradient.
// One thread per (row, column); a row that was never touched (m == v == 0) does not move under dense Adam either.
__global__ void lazy_adam_kernel(float *param, float *exp_avg, float *exp_avg_sq, int *last, long long ld, int H,
                                 const long long *rows, long long R, const float *grad, long long ld_grad, int upto,
                                 AdamHyper h) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = t / H;
    const int c = (int)(t % H);
    if (r >= R) return;
    const long long row = rows ? rows[r] : r;
    const long long o = row * ld + c;
    float p = param[o], m = exp_avg[o], v = exp_avg_sq[o];
    const int from = last[row];
    if (m != 0.f || v != 0.f)
        for (int s = from + 1; s <= upto; s++) adam_one(p, m, v, 0.f, s, h);
    if (grad) adam_one(p, m, v, grad[r * ld_grad + c], upto + 1, h);
    param[o] = p; exp_avg[o] = m; exp_avg_sq[o] = v;
}

__global__ void lazy_adam_mark_kernel(int *last, const long long *rows, long long R, int value) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) last[rows ? rows[r] : r] = value;
}

}  // namespace

extern "C" {

int gp_emb_dropout_fwd(const float *table, int64_t n_table_rows, int64_t ld_table, int32_t H, const int32_t *row_ptr,
                       const int32_t *idx, const float *weight, int64_t B, double p, uint64_t seed, uint64_t offset,
                       float eps, float *out, int64_t ld_out, float *denom_out, void *stream) {
    GpRange nvtx_range("gp_emb_dropout_fwd");
    GP_REQUIRE(B >= 0 && H >= 1 && n_table_rows >= 1, "bad sizes");
    GP_REQUIRE(p >= 0.0 && p < 1.0, "input dropout rate must be in [0,1)");
    if (B == 0) return GP_OK;
    GP_REQUIRE(table && row_ptr && idx && weight && out, "null device buffer");
    const long long blocks = (B * 32 + kEmbBlock - 1) / kEmbBlock;
    GP_REQUIRE(blocks < (1ll << 31), "batch too large for one launch");
    emb_dropout_fwd_kernel<<<(unsigned)blocks, kEmbBlock, 0, (cudaStream_t)stream>>>(
        table, ld_table, H, row_ptr, idx, weight, B, (float)(1.0 / (1.0 - p)), gp_keep_threshold((float)p), seed, offset, eps,
        out, ld_out, denom_out);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

int gp_emb_dropout_bwd(const float *grad_out, int64_t ld_grad_out, int32_t H, const int32_t *row_ptr, const int32_t *slot,
                       const float *weight, const float *denom, int64_t B, double p, uint64_t seed, uint64_t offset,
                       float *grad_rows, int64_t ld_grad_rows, void *stream) {
    GpRange nvtx_range("gp_emb_dropout_bwd");
    GP_REQUIRE(B >= 0 && H >= 1, "bad sizes");
    GP_REQUIRE(p >= 0.0 && p < 1.0, "input dropout rate must be in [0,1)");
    if (B == 0) return GP_OK;
    GP_REQUIRE(grad_out && row_ptr && slot && weight && denom && grad_rows, "null device buffer");
    const long long blocks = (B * 32 + kEmbBlock - 1) / kEmbBlock;
    GP_REQUIRE(blocks < (1ll << 31), "batch too large for one launch");
    emb_dropout_bwd_kernel<<<(unsigned)blocks, kEmbBlock, 0, (cudaStream_t)stream>>>(
        grad_out, ld_grad_out, H, row_ptr, slot, weight, denom, B, (float)(1.0 / (1.0 - p)), gp_keep_threshold((float)p), seed,
        offset, grad_rows, ld_grad_rows);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

int gp_emb_dropout_mask(int64_t nza, int32_t H, double p, uint64_t seed, uint64_t offset, uint8_t *d_mask, void *stream) {
    GP_REQUIRE(nza >= 0 && H >= 1, "bad sizes");
    GP_REQUIRE(p >= 0.0 && p < 1.0, "input dropout rate must be in [0,1)");
    if (nza == 0) return GP_OK;
    GP_REQUIRE(d_mask != nullptr, "null device buffer");
    const long long work = nza * ((H + 3) / 4);
    const long long blocks = std::min<long long>((work + 255) / 256, 148 * 16);
    emb_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(nza, H, gp_keep_threshold((float)p), seed, offset, d_mask);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

int gp_lazy_adam_rows(float *param, float *exp_avg, float *exp_avg_sq, int32_t *last_step, int64_t ld, int32_t H,
                      const int64_t *rows, int64_t R, const float *grad_rows, int64_t ld_grad, int32_t upto, float lr,
                      float beta1, float beta2, float eps, void *stream) {
    GpRange nvtx_range("gp_lazy_adam_rows");
    GP_REQUIRE(R >= 0 && H >= 1 && upto >= 0, "bad sizes");
    if (R == 0) return GP_OK;
    GP_REQUIRE(param && exp_avg && exp_avg_sq && last_step, "null device buffer");
    const long long work = R * H;
    const long long blocks = (work + 255) / 256;
    GP_REQUIRE(blocks < (1ll << 31), "too many rows for one launch");
    AdamHyper h{lr, beta1, beta2, eps};
    lazy_adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, exp_avg, exp_avg_sq, last_step, ld, H,
                                                                          (const long long *)rows, R, grad_rows, ld_grad, upto, h);
    GP_CUDA_TRY(cudaGetLastError());
    lazy_adam_mark_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(last_step, (const long long *)rows, R,
                                                                                          grad_rows ? upto + 1 : upto);
    GP_CUDA_TRY(cudaGetLastError());
    return GP_OK;
}

}  // extern "C"
