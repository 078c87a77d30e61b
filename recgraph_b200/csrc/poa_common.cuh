// Helpers shared by the POA kernels (band formula, run emitter, warp scan).
#pragma once
#include <cuda_runtime.h>

#include "device.h"

namespace rg {

#define NEG_INF (-(1 << 30))
constexpr int WARPS_PER_BLOCK = 8;
constexpr unsigned FULL = 0xffffffffu;

// trace code layout: dir(2) | x(1) | y(1) | dslot(SB) | uslot(SB)
enum { DIR_O = 0, DIR_D = 1, DIR_L = 2, DIR_U = 3 };

__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v = max(v, t);
    }
    return v;
}

// utils.rs:17-72 (simd_version == false)
__device__ __forceinline__ void band_for_row(uint32_t ms, uint32_t me, int32_t r, int32_t L, int32_t bta,
                                             uint32_t& left, uint32_t& right) {
    int32_t t = (r < 0) ? (L + 1 - bta) : (L - r - bta);  // r == -1 encodes usize::MAX (`as i32` == -1)
    int32_t tmp_bs = min((int32_t)ms, t);
    left = tmp_bs < 0 ? 0u : (uint32_t)tmp_bs;
    if (r >= 0 && L > r)
        right = (uint32_t)min(L, max((int32_t)me, L - r) + bta);
    else
        right = (uint32_t)min(L, (int32_t)me + bta);
}

struct RunEmitter {
    rg_run* buf;
    uint32_t cap, n;
    uint32_t op, row, count;
    bool overflow;
    bool ascending;  // rows of graph-consuming ops increase (reverse half of modes 8/9)
    __device__ __forceinline__ void init(rg_run* b, uint32_t c) {
        ascending = false;
        buf = b;
        cap = c;
        n = 0;
        count = 0;
        op = 0;
        row = 0;
        overflow = false;
    }
    __device__ __forceinline__ void flush(int lane) {
        if (count) {
            if (n < cap) {
                if (lane == 0) {
                    buf[n].row = row;
                    buf[n].op_count = (op << 28) | count;
                }
            } else
                overflow = true;
            n++;
            count = 0;
        }
    }
    // `cnt` consecutive steps of the same op starting at row r (rows descending for graph-consuming ops)
    __device__ __forceinline__ void bulk(uint32_t o, uint32_t r, uint32_t cnt, int lane) {
        if (cnt == 0) return;
        bool cont = count && o == op && count + cnt < 0x0fffffffu &&
                    (o == RG_OP_L || o == RG_OP_LPAD ? r == row : (ascending ? r == row + count : r + count == row));
        if (cont)
            count += cnt;
        else {
            flush(lane);
            op = o;
            row = r;
            count = cnt;
        }
    }
    __device__ __forceinline__ void step(uint32_t o, uint32_t r, int lane) {
        bool cont = count && o == op && count < 0x0fffffffu &&
                    (o == RG_OP_L || o == RG_OP_LPAD ? r == row : (ascending ? r == row + count : r + count == row));
        if (cont)
            count++;
        else {
            flush(lane);
            op = o;
            row = r;
            count = 1;
        }
    }
};


}  // namespace rg
