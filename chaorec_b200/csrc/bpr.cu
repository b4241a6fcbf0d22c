// BPR step around the propagation: the loss of
// Model/LightGCN.py:97-135 and its gradient with respect to the propagated table in one kernel.
//
//   loss = -mean_b log(sigmoid(<u,p> - <u,n>) + 1e-5) + reg * (mean(u^2) + mean(p^2) + mean(n^2))
//
// One warp per (user, pos, neg) triple; the three gathered rows live in registers, the two dot
// products are warp-shuffle reductions, the gradient rows are scattered with fp32 atomics into a
// zero-initialised [N, d] table (users / items repeat inside a batch).  The backward of the
// propagation is then the same fused SpMM chain on that table (propagate.py).
#include "common.cuh"

namespace chaorec {
namespace {

template <int VPL>  // values per lane: d = 32 * VPL
__global__ void __launch_bounds__(256) bpr_loss_grad_kernel(const float *__restrict__ emb, const int64_t *__restrict__ users,
                                                            const int64_t *__restrict__ pos, const int64_t *__restrict__ neg,
                                                            int64_t B, int64_t item_offset, float reg, float *__restrict__ loss_terms,
                                                            float *__restrict__ grad) {
    const int64_t b = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = threadIdx.x & 31;
    constexpr int d = 32 * VPL;
    const int64_t ru = users[b], rp = item_offset + pos[b], rn = item_offset + neg[b];
    float u[VPL], p[VPL], n[VPL];
    float sp = 0.f, sn = 0.f, su = 0.f, spp = 0.f, snn = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        u[v] = emb[ru * d + lane + 32 * v];
        p[v] = emb[rp * d + lane + 32 * v];
        n[v] = emb[rn * d + lane + 32 * v];
        sp = fmaf(u[v], p[v], sp);
        sn = fmaf(u[v], n[v], sn);
        su = fmaf(u[v], u[v], su);
        spp = fmaf(p[v], p[v], spp);
        snn = fmaf(n[v], n[v], snn);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
        sn += __shfl_xor_sync(0xffffffffu, sn, o);
        su += __shfl_xor_sync(0xffffffffu, su, o);
        spp += __shfl_xor_sync(0xffffffffu, spp, o);
        snn += __shfl_xor_sync(0xffffffffu, snn, o);
    }
    const float x = sp - sn;
    const float sig = 1.0f / (1.0f + expf(-x));
    const float invB = 1.0f / static_cast<float>(B);
    if (lane == 0) {
        // per-triple share of the loss; the host sums the B terms (deterministic order)
        loss_terms[b] = -logf(sig + 1e-5f) * invB + reg * (su + spp + snn) * invB / static_cast<float>(d);
    }
    // d/dx of -log(sig + 1e-5) = -sig (1 - sig) / (sig + 1e-5)
    const float gx = -(sig * (1.0f - sig)) / (sig + 1e-5f) * invB;
    const float gr = 2.0f * reg * invB / static_cast<float>(d);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        atomicAdd(grad + ru * d + lane + 32 * v, gx * (p[v] - n[v]) + gr * u[v]);
        atomicAdd(grad + rp * d + lane + 32 * v, gx * u[v] + gr * p[v]);
        atomicAdd(grad + rn * d + lane + 32 * v, -gx * u[v] + gr * n[v]);
    }
}

}  // namespace
}  // namespace chaorec

using namespace chaorec;

extern "C" int chaorec_bpr_loss_grad_f32(const float *emb, const int64_t *users, const int64_t *pos_items,
                                         const int64_t *neg_items, int64_t B, int64_t N, int32_t d, int64_t item_offset,
                                         float reg_weight, float *loss_terms, float *grad, void *stream) {
    if (!emb || !users || !pos_items || !neg_items || !loss_terms || !grad || B < 0 || N <= 0) return CHAOREC_E_BADARG;
    if (d != 32 && d != 64 && d != 128) return CHAOREC_E_UNSUPPORTED;
    if (B == 0) return CHAOREC_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = static_cast<unsigned>((B + 7) / 8);
    if (d == 32) bpr_loss_grad_kernel<1><<<blocks, 256, 0, st>>>(emb, users, pos_items, neg_items, B, item_offset, reg_weight, loss_terms, grad);
    else if (d == 64) bpr_loss_grad_kernel<2><<<blocks, 256, 0, st>>>(emb, users, pos_items, neg_items, B, item_offset, reg_weight, loss_terms, grad);
    else bpr_loss_grad_kernel<4><<<blocks, 256, 0, st>>>(emb, users, pos_items, neg_items, B, item_offset, reg_weight, loss_terms, grad);
    return launch_status();
}
