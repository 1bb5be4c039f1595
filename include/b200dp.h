/*
 * b200dp.h -- C ABI of libb200dp.so: the B200 (sm_100a) soft-DP alignment engine.
 *
 * One export per Numba-CUDA kernel of the reference (deepblast/nw_cuda.py,
 * deepblast/sw_cuda.py @ ec661fa), plus the chained adjoint pair, the host-buffer decode,
 * the fused MatrixCrossEntropy and the batched traceback; a reference maintainer binds
 * these with ctypes from the torch.autograd.Functions.  Plain
 * pointers and sizes only; all device memory is owned by the caller (torch's allocator);
 * the library never allocates, frees or retains device memory and never synchronises
 * (b200dp_decode_host owns three internal streams and their events, nothing else).
 * Every launch goes to the CUDA stream passed in (cudaStream_t as void*).
 *
 * Return value: 0 on success, negative for argument errors, positive
 * cudaError_t otherwise; b200dp_last_error() returns a thread-local message.
 *
 * Tensors (all float32, device pointers):
 *   theta, A, ZA   [B, N, M]        contiguous row-major
 *   E, Ed, Ztheta  [B, N+2, M+2]    contiguous row-major (padded lattice, nw.py:347)
 *   Vt, Vtd        [B]
 *   Q, Qd          the reference's [B, N+2, M+2, 3] (nw.py:105) stored STRIP-MAJOR, in the
 *                  order the wavefront produces it, TWO of the three states per cell (x and
 *                  y of deepblast/constants.py:1; the m state is implied: q_m = 1 - q_x - q_y,
 *                  qd_m = -(qd_x + qd_y); q_x = -1 marks a cell whose Q is identically zero):
 *                  lattice cell (i, j) (1-based), stored state c (0 = x, 1 = y) lives at
 *                    b*pair_stride + k*strip_stride + ((j-1) + t)*64 + c*32 + t,
 *                    k = (i-1)/32, t = (i-1)%32,
 *                  with strip_stride = M*64, pair_stride = ceil(N/32)*strip_stride + 31*64
 *                  (b200dp_q_layout()).  Border cells are implicit (zeros, and
 *                  Q[N+1,M+1,:] = 1) and not stored.  The pointer passed is the storage
 *                  base (16-byte aligned); the allocation must be B*pair_stride + pad
 *                  floats.
 *   xlen, ylen     optional int32[B] per-pair lattice sizes (1 <= n <= N,
 *                  1 <= m <= M); NULL = every pair is N x M.  With lengths, each
 *                  pair is computed exactly as the reference computes the slice
 *                  theta[b, :n, :m] on its own (deepblast/alignment.py:165-169);
 *                  E/Ed must then be zero-filled by the caller beforehand.
 *
 * mode: 0 = Needleman-Wunsch (nw.py), 1 = "Smith-Waterman" (sw.py: loops from 2).
 */
#ifndef B200DP_H
#define B200DP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DP_MODE_NW 0
#define B200DP_MODE_SW 1

/* flags (the library reads no environment variables: every dispatch override is a flag) */
#define B200DP_NO_CHAINED    0x1   /* never the chained kernels (large equal-size batches then take the hand-off kernels) */
#define B200DP_NO_TMA        0x2   /* stage tiles with cp.async instead of TMA (debug / unaligned) */
#define B200DP_V1_KERNELS    0x4   /* use the general kernels even where the fast path applies */
#define B200DP_FORCE_CHAINED 0x8   /* the chained kernels at any batch size (tests) */
#define B200DP_WARPS_SHIFT   4     /* bits 4..7: warps per pair (1,2,4,8); 0 = choose automatically */
#define B200DP_CTAS_SHIFT    8     /* bits 8..23: grid size override; 0 = choose automatically */

#define B200DP_TRACEBACK_CPU_RULE  0   /* deepblast/nw.py:401-444 */
#define B200DP_TRACEBACK_CUDA_RULE 1   /* deepblast/nw_cuda.py:273-317 */

int b200dp_version(void);
const char* b200dp_last_error(void);

/* Strip-major Q/Qd geometry for an N x M lattice (all in floats): strips per pair,
 * floats per strip, floats per pair, and the tail padding the allocation needs.
 * Two of the three states of a cell are stored (x, y); the m state is implied: q_m = 1 - q_x - q_y
 * for Q, qd_m = -(qd_x + qd_y) for Qd; q_x = -1 marks a cell whose Q is identically zero
 * (first row / column of the sw.py lattice).
 * As a strided view: Q5[b, k, t, j-1, c] (c = 0: x, 1: y) with strides
 * (pair_stride, strip_stride, 65, 64, 32). */
int b200dp_q_layout(int N, int M, int* K, long long* strip_stride, long long* pair_stride,
                    long long* pad);

/* replaces _forward_pass_kernel, deepblast/nw_cuda.py:46-79 (sw_cuda.py:46-79):
 * theta, A -> Vt, Q (lattice cells; the zero borders are implicit). */
int b200dp_fwd(const float* theta, const float* A, float* Q, float* Vt,
               const int32_t* xlen, const int32_t* ylen, int B, int N, int M,
               int mode, int flags, void* stream);

/* replaces _backward_pass_kernel, deepblast/nw_cuda.py:82-102 (sw_cuda.py:82-102):
 * Et (element stride et_stride, 0 for an expanded scalar), Q -> E, borders included
 * (zeros, E[N+1, M+1] = Et). */
int b200dp_bwd(const float* Et, long long et_stride, const float* Q, float* E,
               const int32_t* xlen, const int32_t* ylen, int B, int N, int M,
               int mode, int flags, void* stream);

/* b200dp_bwd that, where the chained kernel takes the batch, also writes the interior of E as
 * a contiguous [B, N, M] tensor Ei (what b200dp_adj_fwd3 reads if a double backward follows);
 * *wrote_ei = 1 if it did, 0 if Ei was left untouched (the caller then slices E itself). */
int b200dp_bwd_keep_interior(const float* Et, long long et_stride, const float* Q, float* E,
                             float* Ei, int* wrote_ei, int B, int N, int M, int mode, int flags,
                             void* stream);

/* replaces _adjoint_forward_pass_kernel, deepblast/nw_cuda.py:105-139:
 * Q, Ztheta (padded), ZA -> Vtd, Qd. */
int b200dp_adj_fwd(const float* Q, const float* Ztheta, const float* ZA, float* Vtd,
                   float* Qd, const int32_t* xlen, const int32_t* ylen, int B, int N,
                   int M, int flags, void* stream);

/* replaces _adjoint_backward_pass_kernel, deepblast/nw_cuda.py:142-165:
 * E, Q, Qd -> Ed (padded, zero borders). */
int b200dp_adj_bwd(const float* E, const float* Q, const float* Qd, float* Ed,
                   const int32_t* xlen, const int32_t* ylen, int B, int N, int M,
                   int flags, void* stream);

/* The same pair of sweeps for large batches of equal-size lattices (the double backward of a
 * training step, trainer.py:154-171 -> nw_cuda.py:243-262), on the chained kernels:
 *   b200dp_adj3_applicable(B, N, M) -> 1 if they take the shape (else use the pair above);
 *   b200dp_adj_fwd3: Q, Zt = the INTERIOR of Ztheta as a contiguous [B, N, M] tensor, ZA
 *     [B, N, M] or NULL (= zeros, the usual case), E interior [B, N, M] or NULL
 *     -> Vtd [B] and the strip-major stream QdE = Qd * E (Qd itself when E is NULL);
 *   b200dp_adj_bwd3: Q, QdE -> Ed [B, N+2, M+2] (zero borders) and / or its interior as a
 *     contiguous [B, N, M] tensor (either pointer may be NULL, not both). */
int b200dp_adj3_applicable(int B, int N, int M);
int b200dp_adj_fwd3(const float* Q, const float* Zt, const float* ZA, const float* E, float* Vtd,
                    float* QdE, int B, int N, int M, int flags, void* stream);
int b200dp_adj_bwd3(const float* Q, const float* QdE, float* Ed, float* Ed_interior, int B, int N,
                    int M, int flags, void* stream);

/* replaces the Python walk NeedlemanWunschDecoder.traceback,
 * deepblast/nw.py:401-444 (variant 0) / deepblast/nw_cuda.py:273-317 (variant 1),
 * batched: grad [B, N, M] with element strides (sb, si, sj) -> out [B, cap, 3]
 * int32 triples (i, j, state) in the reference's returned order, len [B]
 * (-1: cap too small, -2: the reference would raise IndexError). */
int b200dp_traceback(const float* grad, long long sb, long long si, long long sj,
                     const int32_t* xlen, const int32_t* ylen, int B, int N, int M,
                     int variant, int32_t* out, int cap, int32_t* len, void* stream);

/* replaces the per-pair Python loop of MatrixCrossEntropy.__call__, deepblast/losses.py:9-48
 * (trainer.py:154-171), the step right after decode: Ytrue, G [B, N, M] contiguous (G NULL =
 * all counted), Ypred [B, N, M] with element strides (pb, pi, 1) -- e.g. the padded E in
 * place -- and optional per-pair lengths -> pair_loss[b] = l_b / B (loss = sum_b) and
 * pair_count[b]; the backward writes dloss/dYpred [B, N, M] (zeros outside the mask)
 * scaled by the scalar gout[0]. */
int b200dp_mxent_fwd(const float* Ytrue, const float* Ypred, long long pb, long long pi,
                     const float* G, const int32_t* xlen, const int32_t* ylen, int B, int N,
                     int M, float* pair_loss, float* pair_count, void* stream);
int b200dp_mxent_bwd(const float* Ytrue, const float* Ypred, long long pb, long long pi,
                     const float* G, const int32_t* xlen, const int32_t* ylen, int B, int N,
                     int M, const float* pair_count, const float* gout, float* grad,
                     void* stream);

/* Host-buffer form of NeedlemanWunschDecoder.decode (deepblast/nw_cuda.py:319-325: forward,
 * then autograd.grad of sum(Vt), i.e. _forward_pass_kernel + _backward_pass_kernel) for a
 * caller whose theta / A live in HOST memory (pinned for full PCIe speed):
 *   theta_h, A_h [B, N, M], Et_h [B] or NULL (= ones, what sum(Vt).backward() feeds)
 *   -> Vt_h [B], E_h [B, N+2, M+2] (padded as the reference's E; dVt/dtheta = E[:, 1:-1, 1:-1]).
 * The batch is cut into chunks of chunk_pairs that flow upload -> fwd -> bwd -> download
 * through three slots of the caller-provided DEVICE workspace on three internal streams, so
 * uploads, sweeps and downloads overlap.  Asynchronous: forked from `stream` and joined back
 * into it; the host buffers are valid once the caller has synchronised `stream`. */
size_t b200dp_decode_host_workspace(int N, int M, int chunk_pairs);
int b200dp_decode_host(const float* theta_h, const float* A_h, const float* Et_h, float* Vt_h,
                       float* E_h, int B, int N, int M, int mode, int chunk_pairs,
                       void* workspace, size_t workspace_bytes, int flags, void* stream);

/* Inference from HOST buffers: what DeepBLAST.align needs per pair is the alignment PATH
 * (deepblast/trainer.py:80-88 -> alignment.py:160-171 -> the walk of nw_cuda.py:273-317).  The same
 * chunked pipeline as b200dp_decode_host, with the walk run on the device (b200dp_traceback on each
 * chunk's E) so that only the paths and scores cross PCIe:
 *   -> Vt_h [B], paths_h [B, cap, 3] int32 triples (i, j, state), cap = b200dp_align_host_path_cap(N, M),
 *      len_h [B] (steps of each path; -2 / -1 as b200dp_traceback), and, when E_h is not NULL, the
 *      padded E as well.  variant: B200DP_TRACEBACK_*_RULE. */
size_t b200dp_align_host_workspace(int N, int M, int chunk_pairs);
int b200dp_align_host_path_cap(int N, int M);
int b200dp_align_host(const float* theta_h, const float* A_h, float* Vt_h, int32_t* paths_h,
                      int32_t* len_h, float* E_h, int B, int N, int M, int mode, int variant,
                      int chunk_pairs, void* workspace, size_t workspace_bytes, int flags,
                      void* stream);

/* ---- strip-queue family: batches of any shape -- ragged (per-pair lengths), PACKED, small
 * batches of long pairs, large batches of equal pairs (softdp_sq.cuh).  Same passes as above
 * (one export per reference kernel, deepblast/nw_cuda.py:46-165); with per-pair lengths each
 * pair is computed exactly as the reference computes the slice theta[b, :n_b, :m_b] on its own
 * (deepblast/alignment.py:165-169).
 *
 * b200dp_plan_build (HOST code, no CUDA call) cuts the pairs into strips of 32 rows and writes
 * the two work-queue tables (forward order, backward order; 64 bytes per strip) that the caller
 * uploads once per batch and reuses for all four sweeps.  Operand layouts:
 *   dense  (packed = 0): theta / A / E / Ztheta / Ed are contiguous [B, N, M] (M % 4 == 0), pair
 *                        b uses the top-left n_b x m_b corner; E / Ed are the INTERIOR of the
 *                        reference's padded tensors (E_ref[:, 1:-1, 1:-1], nw.py:339): no
 *                        border is stored, the kernels touch nothing outside a pair's corner;
 *   packed (packed = 1): one flat buffer, pair b is an n_b x pitch_b row-major block at
 *                        element offset pair_off[b], pitch_b = (m_b + 3) & ~3 -- the layout of
 *                        dataset/utils.py:214-251 carried one step further, no padding to the
 *                        global maximum length at all.
 * Q / QdE: the strip-major layout described at the top with the pair's own m_b, pair b at float
 * offset q_off[b] of a buffer of info.q_floats floats.
 * Lengths outside [0, N] x [0, M] are clamped like the reference's slices; a pair with n_b = 0
 * or m_b = 0 has no strips (its Vt is not written: zero-fill Vt beforehand).
 * Call with fwd_tab = bwd_tab = NULL to size the buffers (info), then again to fill them.
 * warps_fwd / warps_bwd: b200dp_sq_resident_warps() of the forward / backward sweep -- the ticket
 * order is a list schedule for that many warps (longest remaining dependency chain first, a
 * strip not before its predecessor is far enough ahead); <= 0 picks a default.
 *
 * The launchers take a DEVICE workspace of b200dp_sq_workspace_bytes(info.bnd_words) bytes,
 * zero-filled ONCE when allocated and used by one launch at a time (launches on one stream are
 * fine): the kernels leave the ticket counter clean and count the launches in the workspace
 * itself (the tag that tells this launch's hand-off words from stale ones), so launches may be
 * captured in CUDA graphs and replayed.  flags: B200DP_CTAS_SHIFT (grid override), B200DP_SQ_RING_SHIFT. */
#define B200DP_SQ_RING_SHIFT 24    /* bits 24..27: tile ring depth override (all fast kernels); 0 = default */
#define B200DP_SQ_DBG_SHIFT  28    /* bits 28..30: diagnostics (timing experiments; results are wrong when set) */

typedef struct b200dp_plan_info {
    int nstrips;               /* records per table */
    int max_m;
    long long q_floats;        /* floats of a Q / QdE buffer */
    long long bnd_words;       /* 8-byte words of boundary scratch (b200dp_sq_workspace_bytes) */
    long long packed_floats;   /* floats of a theta / A / E buffer in this layout */
    long long cells;           /* sum of n_b * m_b */
    int grid_fwd, grid_bwd;    /* warps (CTAs) to launch the forward- / backward-direction sweeps with
                                  (pass through the B200DP_CTAS_SHIFT field of flags) */
} b200dp_plan_info;

int b200dp_plan_build(const int32_t* xlen, const int32_t* ylen, int B, int N, int M, int packed,
                      int warps_fwd, int warps_bwd, b200dp_plan_info* info, long long* pair_off,
                      long long* q_off, void* fwd_tab, void* bwd_tab, int tab_capacity);
size_t b200dp_sq_workspace_bytes(long long bnd_words);
/* resident warps of sweep `kind` (0 fwd, 1 bwd, 2 adjoint fwd, 3 adjoint bwd) on the current device */
int b200dp_sq_resident_warps(int kind);

/* diagnostics: the launches of the calling thread record {start, end} (globaltimer ns) of every
 * ticket into trace[2 * nstrips] (device memory); NULL switches it off again */
void b200dp_sq_set_trace(void* trace);

/* _forward_pass_kernel (nw_cuda.py:46-79).  Q = NULL: score only, Vt alone
 * (deepblast/alignment.py:127-137 calls ddp(theta, A) under no_grad). */
int b200dp_sq_fwd(const void* fwd_tab, int nstrips, void* workspace,
                  const float* theta, const float* A, float* Q, float* Vt, int mode, int flags,
                  void* stream);
/* The same sweep for a plan in the DENSE layout (theta, A contiguous [B, N, M], M % 4 == 0): the operand
 * tiles travel as 16 x 16 TMA boxes issued by one elected lane instead of per-lane 16-byte copies. */
int b200dp_sq_fwd_dense(const void* fwd_tab, int nstrips, void* workspace,
                        const float* theta, const float* A, float* Q, float* Vt,
                        int B, int N, int M, int mode, int flags, void* stream);
/* _backward_pass_kernel (nw_cuda.py:82-102): Et, Q -> E (interior layout). */
int b200dp_sq_bwd(const void* bwd_tab, int nstrips, void* workspace,
                  const float* Et, long long et_stride, const float* Q, float* E, int mode,
                  int flags, void* stream);
/* _adjoint_forward_pass_kernel (nw_cuda.py:105-139): Q, Zt (interior layout), ZA or NULL, E
 * (interior layout) or NULL -> Vtd, QdE = Qd * E (Qd itself when E is NULL). */
int b200dp_sq_adj_fwd(const void* fwd_tab, int nstrips, void* workspace,
                      const float* Q, const float* Zt, const float* ZA, const float* E,
                      float* Vtd, float* QdE, int flags, void* stream);
/* _adjoint_backward_pass_kernel (nw_cuda.py:142-165): Q, QdE -> Ed (interior layout). */
int b200dp_sq_adj_bwd(const void* bwd_tab, int nstrips, void* workspace,
                      const float* Q, const float* QdE, float* Ed, int flags, void* stream);

/* ---- cluster kernels: SMALL batches of LONG equal-size lattices (what the reference trains and infers
 * on: a few dozen pairs of up to 1024 x 1024, deepblast/trainer.py:375; one pair at a time,
 * alignment.py:165-169).  One thread-block cluster per pair, its strips dealt round-robin to the
 * cluster's CTAs, boundary rows handed over through distributed shared memory (softdp_cl.cuh).
 * theta / A dense [B, N, M] (M % 4 == 0, N > 32); Q as everywhere (NULL: score only); E is the INTERIOR
 * [B, N, M] like the strip-queue kernels'.  b200dp_cl_applicable: 0 when the strip-queue kernels should
 * take the batch (it is not bound by one pair's dependency chain), else the cluster size.
 * flags: bits 4..7 force the cluster size (1, 2, 4, 8). */
int b200dp_cl_applicable(int B, int N, int M);
int b200dp_cl_fwd(const float* theta, const float* A, float* Q, float* Vt, int B, int N, int M,
                  int mode, int flags, void* stream);
int b200dp_cl_bwd(const float* Et, long long et_stride, const float* Q, float* E, int B, int N, int M,
                  int mode, int flags, void* stream);

/* ---- the step before the DP: theta = softplus(zx zy^T), A = logsigmoid(gx gy^T)
 * (deepblast/alignment.py:122-123,134-135,162-163) as one batched tcgen05 GEMM launch with the
 * activation fused into the epilogue (softdp_gemm.cu).  zx, gx [B, Lx, D], zy, gy [B, Ly, D] fp32
 * contiguous, D % 64 == 0; fp32 accuracy through a bf16 hi/lo split (three tensor-core passes).
 * Outputs in the DP's operand layout: pair_off == NULL: dense [B, Lx, Ly]; else pair b at element
 * offset pair_off[b] with pitch (m_b + 3) & ~3 (b200dp_plan_build, packed).  xlen / ylen (device
 * int32[B], nullable): only the n_b x m_b corner of pair b is computed and written.
 * workspace: DEVICE memory, b200dp_theta_a_workspace() bytes (the bf16 copies of the embeddings). */
size_t b200dp_theta_a_workspace(int B, int Lx, int Ly, int D);
int b200dp_theta_a(const float* zx, const float* zy, const float* gx, const float* gy, int B, int Lx,
                   int Ly, int D, const int32_t* xlen, const int32_t* ylen, const long long* pair_off,
                   float* theta, float* A, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DP_H */
