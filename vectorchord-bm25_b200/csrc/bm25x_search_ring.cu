// bm25x_search_ring.cu — instantiations of k_search_ring (bm25x_search_ring.cuh) for one pool size.
// Compiled once per pool capacity (-DBM25X_RING_KP=64|256|2048) so that the term-count classes build in parallel.
#include "bm25x_common.h"

#include "bm25x_search_ring.cuh"

#ifndef BM25X_RING_KP
#error "compile with -DBM25X_RING_KP=<pool capacity>"
#endif

namespace {

template <class C>
int launch_ring(int device, int sm_count, const SearchParams &sp_in, cudaStream_t stream) {
    static bool configured[64] = {false};
    auto kern = k_search_ring<C>;
    if (!configured[device & 63]) {
        BM25X_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::total));
        configured[device & 63] = true;
    }
    uint32_t grid = (uint32_t)sm_count;  // persistent: one CTA per SM, every warp pulls queries from the work counter
    const uint32_t need = (sp_in.nq + C::WARPS - 1) / C::WARPS;
    if (grid > need) grid = need;
    SearchParams sp = sp_in;
    sp.pool_scratch = nullptr;
    if (C::POOL_GLOBAL) {  // stream-ordered scratch for the per-warp pools, released right behind the kernel
        const size_t bytes = (size_t)grid * C::WARPS * (size_t)C::KP * 16;
        BM25X_CUDA_TRY(cudaMallocAsync((void **)&sp.pool_scratch, bytes, stream));
    }
    kern<<<grid, C::THREADS, C::total, stream>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (sp.pool_scratch) cudaFreeAsync(sp.pool_scratch, stream);
    BM25X_CUDA_TRY(e);
    return BM25X_OK;
}

}  // namespace

#define BM25X_RING_ENTRY2(kp) bm25x_launch_ring_kp##kp
#define BM25X_RING_ENTRY(kp) BM25X_RING_ENTRY2(kp)

// M = term-count class of the launch (1, 2, 3, 4, 8, 16, 32); phase = RCfg::PH (1 / 2 only for 2..4 terms, KP <= 256)
int BM25X_RING_ENTRY(BM25X_RING_KP)(int device, int sm_count, const SearchParams &sp, int M, int phase, cudaStream_t stream) {
#if BM25X_RING_KP <= 256
    if (phase == 1) switch (M) {
            case 2: return launch_ring<RCfg<2, BM25X_RING_KP, 1>>(device, sm_count, sp, stream);
            case 3: return launch_ring<RCfg<3, BM25X_RING_KP, 1>>(device, sm_count, sp, stream);
            case 4: return launch_ring<RCfg<4, BM25X_RING_KP, 1>>(device, sm_count, sp, stream);
            default: break;
        }
    if (phase == 2) switch (M) {
            case 2: return launch_ring<RCfg<2, BM25X_RING_KP, 2>>(device, sm_count, sp, stream);
            case 3: return launch_ring<RCfg<3, BM25X_RING_KP, 2>>(device, sm_count, sp, stream);
            case 4: return launch_ring<RCfg<4, BM25X_RING_KP, 2>>(device, sm_count, sp, stream);
            default: break;
        }
    if (phase == 3) switch (M) {
            case 2: return launch_ring<RCfg<2, BM25X_RING_KP, 3>>(device, sm_count, sp, stream);
            case 3: return launch_ring<RCfg<3, BM25X_RING_KP, 3>>(device, sm_count, sp, stream);
            case 4: return launch_ring<RCfg<4, BM25X_RING_KP, 3>>(device, sm_count, sp, stream);
            case 8: return launch_ring<RCfg<8, BM25X_RING_KP, 3>>(device, sm_count, sp, stream);
            default: break;
        }
    if (phase == 4) switch (M) {
            case 2: return launch_ring<RCfg<2, BM25X_RING_KP, 4>>(device, sm_count, sp, stream);
            case 3: return launch_ring<RCfg<3, BM25X_RING_KP, 4>>(device, sm_count, sp, stream);
            case 4: return launch_ring<RCfg<4, BM25X_RING_KP, 4>>(device, sm_count, sp, stream);
            case 8: return launch_ring<RCfg<8, BM25X_RING_KP, 4>>(device, sm_count, sp, stream);
            default: break;
        }
#endif
    if (phase != 0) {
        bm25x_set_error("k_search_ring: no two-phase launch for %d terms / pool %d", M, (int)BM25X_RING_KP);
        return BM25X_ERR_INVALID;
    }
    switch (M) {
        case 1: return launch_ring<RCfg<1, BM25X_RING_KP>>(device, sm_count, sp, stream);
        case 2: return launch_ring<RCfg<2, BM25X_RING_KP>>(device, sm_count, sp, stream);
        case 3: return launch_ring<RCfg<3, BM25X_RING_KP>>(device, sm_count, sp, stream);
        case 4: return launch_ring<RCfg<4, BM25X_RING_KP>>(device, sm_count, sp, stream);
        case 8: return launch_ring<RCfg<8, BM25X_RING_KP>>(device, sm_count, sp, stream);
        case 16: return launch_ring<RCfg<16, BM25X_RING_KP>>(device, sm_count, sp, stream);
        default: return launch_ring<RCfg<32, BM25X_RING_KP>>(device, sm_count, sp, stream);
    }
}
