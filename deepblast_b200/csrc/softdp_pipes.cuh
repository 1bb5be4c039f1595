// softdp_pipes.cuh -- per-warp shared-memory tile pipelines fed by TMA.
//
// Every warp owns its rings and its mbarriers; nothing here synchronises the CTA.
// A pipe streams the tiles of the warp's strips in ONE linear order that runs
// across strip and pair boundaries, so the prefetch never drains between strips.
#pragma once
#include "softdp_common.cuh"

namespace b200dp {

// Bookkeeping of one ring of R slots.  AHEAD = how far past the tile being waited
// on the producer may run: R-1 for tiles that die when the next one starts,
// R-2 for the skewed row-major tiles (tile a-1 is still read while tile a is live).
template <int R, int AHEAD>
struct TilePipe {
    int issued;        // tiles issued so far, counted from tile 0 of the CURRENT strip
    unsigned islot;    // slot the next issue goes to
    unsigned wslot;    // slot the next wait looks at
    unsigned phases;   // one parity bit per slot

    __device__ __forceinline__ void reset() {
        issued = 0;
        islot = wslot = 0;
        phases = 0;
    }
    // Issue everything the ring can hold when the warp is about to consume tile `a`
    // of the current strip (Tcur tiles); tiles past the strip's end come from the
    // warp's next strip (Tnxt tiles, if nvalid).  fn(from_next, tile_index, slot).
    template <class IssueFn>
    __device__ __forceinline__ void pump(int a, int Tcur, bool nvalid, int Tnxt, IssueFn&& fn) {
        while (issued <= a + AHEAD) {
            if (issued < Tcur) fn(false, issued, islot);
            else if (nvalid && issued - Tcur < Tnxt) fn(true, issued - Tcur, islot);
            else break;
            issued++;
            islot = (islot + 1 == R) ? 0u : islot + 1;
        }
    }
    // Wait for the next tile in linear order; returns its slot.
    __device__ __forceinline__ unsigned wait(uint64_t* bars) {
        mbar_wait(&bars[wslot], (phases >> wslot) & 1u);
        phases ^= 1u << wslot;
        unsigned s = wslot;
        wslot = (wslot + 1 == R) ? 0u : wslot + 1;
        return s;
    }
    __device__ __forceinline__ void next_strip(int Tcur) { issued -= Tcur; }
};

// ---- row-major 32x32 tile (theta, A, ZA, Ztheta, E) ---------------------------
// Logical element (r, c) of a pair (0-based lattice row/col) lives at
// base + r*pitch + c.  Tile (rb, cb) covers rows 32rb.., cols 32cb..; it lands in
// shared memory dense [32][32] so lane t reading (row t, col c) hits bank c%32:
// the wavefront's skewed read (c = s - t) is conflict-free.
struct RowSrc {
    const float* base;   // element (0,0) of pair 0
    long long pair_stride;
    int pitch;
    int rows, cols;      // addressable extent per pair (reads beyond are zero-filled)
};

// Generic path: 32 lanes x 32 cp.async (4 B).  The CALLER arrives on the slot's
// mbarrier exactly once per lane per slot (cp_async_mbar_arrive_noinc) after all the
// tensors that share the slot have been issued.
__device__ __forceinline__ void row_tile_load_generic(float* dst, const RowSrc& src, int pair, int rb,
                                                      int cb, int lane) {
    const float* pb = src.base + (long long)pair * src.pair_stride;
    const int col = cb * kTile + lane;
    const bool cok = col >= 0 && col < src.cols;
#pragma unroll 8
    for (int r = 0; r < kTile; ++r) {
        const int row = rb * kTile + r;
        const bool ok = cok && row < src.rows;
        const float* g = ok ? (pb + (long long)row * src.pitch + col) : src.base;
        cp_async4_zfill(dst + r * kTile + lane, g, ok);
    }
}

// ---- strip-major Q tile: kDiagRows consecutive wavefront steps = 4 KB contiguous -----
// Loads steps [sig_lo, sig_lo + 16) of a strip into dst[16][2][32]; steps below 0 (the
// last tile of a right-to-left sweep) are skipped and their slots left untouched.
// kTMA: lane 0 arms the mbarrier (for expect_copies equal copies) and issues one 1-D bulk copy.  Otherwise every lane
// issues 16-byte cp.async and the CALLER arrives (cp_async_mbar_arrive_noinc).
template <bool kTMA>
__device__ __forceinline__ void q_tile_load(float* dst, uint64_t* bar, const float* strip, int sig_lo, int lane,
                                            int expect_copies = 1) {
    const int lo = sig_lo < 0 ? 0 : sig_lo;
    const int nsteps = kDiagRows - (lo - sig_lo);
    const float* src = strip + (long long)lo * kStepFloats;
    float* d = dst + (lo - sig_lo) * kStepFloats;
    if (kTMA) {
        if (elect_one()) {
            // expect_copies same-sized copies complete on this barrier; the first one arms it
            if (expect_copies > 0) mbar_expect_tx(bar, (uint32_t)(expect_copies * nsteps) * kStepFloats * 4);
            tma_bulk_load(d, src, (uint32_t)nsteps * kStepFloats * 4, bar);
        }
    } else {
        for (int c = lane; c < nsteps * (kStepFloats / 4); c += 32) cp_async16(d + c * 4, src + c * 4);
    }
}

}  // namespace b200dp
