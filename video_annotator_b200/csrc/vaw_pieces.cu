// vaw_pieces.cu -- builds the per-piece polynomial tables of the map (see vaw_pieces.cuh).
//
// One CTA handles up to 32 horizontally adjacent pieces of one piece-row of one frame:
//   1. anchors: the projection of /root/reference/opencv/createMap.cl:15-49 evaluated in
//      double precision on the shared node grid (6 x 4 nodes per piece, edges shared);
//   2. Lagrange -> monomial along u, then along v (double);
//   3. base = nearest integer of the centre coefficient; offsets stored as fp32;
//   4. certificates: regularity (q.z range, optical axis not inside), accuracy against the
//      exact projection at three interior check points, and the coordinate range
//      (|offset| <= L1 norm of the non-constant coefficients) that classifies the piece as
//      interior / outside / mixed with respect to the source frame.
// This is a few hundred double-precision instructions per 4096 output pixels.
#include <math.h>
#include "vaw_pieces.cuh"
#include "vaw_project64.cuh"

namespace vaw {

namespace {

#ifndef VAW_BUILDER_CHUNK
#define VAW_BUILDER_CHUNK 32
#endif
constexpr int kChunk = VAW_BUILDER_CHUNK;        // pieces per CTA along u
constexpr int kNodesU = kChunk * kDegU + 1;      // 161 shared nodes
#ifndef VAW_BUILDER_THREADS
#define VAW_BUILDER_THREADS 128  // small CTAs: the kernel is latency-bound (C3, 64 frames: 65 us at 128, 69 at 160, 76 at 192, 80 at 256 and at 64)
#endif
constexpr int kThreads = VAW_BUILDER_THREADS;

struct RotF { float r[9]; };

// Anchor node positions: 128 (ig / 5) + 128 (ig % 5) / 5 and ph (jg / 3) + ph (jg % 3) / 3.  The quotients are
// written as correctly rounded constants (ph is a power of two, so scaling commutes with the rounding): the
// same doubles the divisions produce, without a double-precision division subroutine per anchor (a zero
// numerator sent every fifth / third one down its slow path; with 1 / q2 and atan() that was 16 % of the
// builder's instructions).
__device__ __forceinline__ double node_u(int ig)
{
    const int m = ig % kDegU;
    const double f = m == 0 ? 0.0 : (m == 1 ? 128.0 / 5 : (m == 2 ? 256.0 / 5 : (m == 3 ? 384.0 / 5 : 512.0 / 5)));
    return 128.0 * (ig / kDegU) + f;
}
__device__ __forceinline__ double node_v(int jg, int ph)
{
    const int m = jg % kDegV;
    const double f = m == 0 ? 0.0 : (m == 1 ? 1.0 / 3 : 2.0 / 3);
    return (double)ph * (jg / kDegV) + (double)ph * f;
}
static_assert(kDegU == 5 && kDegV == 3, "node_u / node_v spell out the fifths and thirds");

__device__ __forceinline__ float coef(const PieceRec& r, int c, int i, int j) { return c ? r.c[i][j].y : r.c[i][j].x; }
__device__ __forceinline__ float base_of(const PieceRec& r, int c) { return c ? r.base.y : r.base.x; }

__device__ __forceinline__ double poly_eval(const PieceRec& r, int c, double s, double t)
{
    double acc = 0.0;
#pragma unroll
    for (int i = kDegU; i >= 0; --i) {
        double a = 0.0;
#pragma unroll
        for (int j = kDegV; j >= 0; --j) a = a * t + (double)coef(r, c, i, j);
        acc = acc * s + a;
    }
    return (double)base_of(r, c) + acc;
}

__global__ void __launch_bounds__(kThreads)
build_pieces_kernel(const GeomD g, const PieceBasis basis, const float* __restrict__ rots, const RotF rot0,
                    PieceRec* __restrict__ table)
{
    __shared__ double anchors[kNv][kNodesU][2];
    __shared__ double ucoef[kChunk][2][kNv][kNu];
    __shared__ PieceRec recs[kChunk];
    __shared__ int bad[kChunk];

    const int ph = g.piece_h;
    const int npx = pieces_x(g.out_w), npy = pieces_y(g.out_h, ph);
    const int p0 = blockIdx.x * kChunk, py = blockIdx.y, frame = blockIdx.z;
    const int np = min(kChunk, npx - p0);
    const int tid = threadIdx.x;

    RotD R;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.r[i] = (double)(rots ? __ldg(rots + (size_t)frame * 9 + i) : rot0.r[i]);

    if (tid < kChunk) bad[tid] = 0;
    __syncthreads();

    // 1. anchors on the shared node grid
    const int nodes = np * kDegU + 1;
#ifndef VAW_BUILDER_UNROLL
#define VAW_BUILDER_UNROLL 1  // anchors per thread evaluated side by side (instruction-level parallelism of the fp64 chains)
#endif
    constexpr int kAnchorUnroll = VAW_BUILDER_UNROLL;
    int b = 0, ig = tid;  // node (row b, column ig) of flat index idx = b * nodes + ig, kept without a division per anchor
    while (ig >= nodes) { ig -= nodes; ++b; }
#pragma unroll kAnchorUnroll
    for (int idx = tid; idx < nodes * kNv; idx += kThreads) {
        const Ray a = project(g, R, node_u(p0 * kDegU + ig), node_v(py * kDegV + b, ph));
        anchors[b][ig][0] = a.mx;
        anchors[b][ig][1] = a.my;
        if (!(isfinite(a.mx) && isfinite(a.my))) {
            if (ig / kDegU < np) bad[ig / kDegU] = 1;
            if (ig % kDegU == 0 && ig > 0) bad[ig / kDegU - 1] = 1;
        }
        ig += kThreads;
        while (ig >= nodes) { ig -= nodes; ++b; }
    }
    __syncthreads();

    // 2a. along u: (piece, coordinate, v-node) -> 6 monomial coefficients in s
    for (int task = tid; task < np * 2 * kNv; task += kThreads) {
        const int p = task / (2 * kNv), c = (task / kNv) & 1, b = task % kNv;
        double f[kNu];
#pragma unroll
        for (int a = 0; a < kNu; ++a) f[a] = anchors[b][p * kDegU + a][c];
#pragma unroll
        for (int i = 0; i < kNu; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int a = 0; a < kNu; ++a) acc += basis.mu[i][a] * f[a];
            ucoef[p][c][b][i] = acc;
        }
    }
    __syncthreads();

    // 2b. along v: (piece, coordinate, power of s) -> 4 monomial coefficients in t; 3. base
    for (int task = tid; task < np * 2 * kNu; task += kThreads) {
        const int p = task / (2 * kNu), c = (task / kNu) & 1, i = task % kNu;
        double co[kNv];
#pragma unroll
        for (int j = 0; j < kNv; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < kNv; ++b) acc += basis.mv[j][b] * ucoef[p][c][b][i];
            co[j] = acc;
        }
        if (i == 0) {
            const double base = nearbyint(co[0]);
            (c ? recs[p].base.y : recs[p].base.x) = (float)base;
            co[0] -= base;
        }
#pragma unroll
        for (int j = 0; j < kNv; ++j) (c ? recs[p].c[i][j].y : recs[p].c[i][j].x) = (float)co[j];
    }
    __syncthreads();

    // 4. certificates, all in one phase on disjoint thread ranges:
    //    tasks [0, 3 np): accuracy -- exact projection against the fp32 polynomial where the
    //    interpolation error of equispaced nodes peaks (inside the first / last node intervals);
    //    tasks [3 np, 4 np): regularity, coordinate range -> flags and the source box.
    for (int task = tid; task < np * 4; task += kThreads) {
        if (task < np * 3) {
            const int p = task / 3, which = task % 3;
            const double du = which == 0 ? 9.0 : (which == 1 ? 119.0 : 70.0);
            const double dv = (which == 1 ? 0.894 : 0.106) * ph;
            const Ray e = project(g, R, 128.0 * (p0 + p) + du, (double)ph * py + dv);
            const double s = (du - 63.5) / 64.0, t = (dv - 0.5 * (ph - 1)) * (2.0 / ph);
            const double ex = poly_eval(recs[p], 0, s, t) - e.mx;
            const double ey = poly_eval(recs[p], 1, s, t) - e.my;
            if (!(fabs(ex) <= 5e-5 && fabs(ey) <= 5e-5)) bad[p] = 1;
            continue;
        }
        const int p = task - np * 3;
        bool ok = true;
        bool pos0 = true, neg0 = true, pos1 = true, neg1 = true;
#pragma unroll
        for (int corner = 0; corner < 4; ++corner) {
            const Ray a = ray_only(g, R, 128.0 * (p0 + p) + ((corner & 1) ? 128.0 : 0.0),
                                   (double)ph * py + ((corner & 2) ? (double)ph : 0.0));
            ok = ok && a.q2 >= 0.015625 && a.q2 <= 64.0 && fabs(a.q0) <= 64.0 && fabs(a.q1) <= 64.0;
            const double eps = 9.5367431640625e-07;  // 2^-20
            pos0 = pos0 && a.q0 >= eps; neg0 = neg0 && a.q0 <= -eps;
            pos1 = pos1 && a.q1 >= eps; neg1 = neg1 && a.q1 <= -eps;
        }
        // optical axis not inside: the reference yields NaN at r == 0 (createMap.cl:38-39), which only the per-pixel
        // path reproduces; the other projection pairs are analytic there
        if (g.projection == 0) ok = ok && (pos0 || neg0 || pos1 || neg1);
        PieceRec& rec = recs[p];
        // coordinate range: |offset| <= L1 norm of the non-constant coefficients (|s|, |t| <= 1)
        double lo[2], hi[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double rad = 0.0;
#pragma unroll
            for (int i = 0; i < kNu; ++i)
#pragma unroll
                for (int j = 0; j < kNv; ++j)
                    if (i | j) rad += fabs((double)coef(rec, c, i, j));
            const double centre = (double)base_of(rec, c) + (double)coef(rec, c, 0, 0);
            lo[c] = centre - rad - 0.01;
            hi[c] = centre + rad + 0.01;
        }
        ok = ok && fabs(lo[0]) < 30000.0 && fabs(hi[0]) < 30000.0 && fabs(lo[1]) < 30000.0 && fabs(hi[1]) < 30000.0;
        uint32_t flags = 0;
        PieceBox box = {0, -1, 0, -1, 0, -1, 0, -1};
        if (ok) {
            flags |= kPiecePoly;
            const double W = g.src_w, H = g.src_h;
            // luma taps ix, ix+1 in [0, W-1]  <=>  0 <= m < W-1;  chroma: 0.5 <= mean < W-1.5.  With a halo of h taps per
            // side (INTER_CUBIC 1, INTER_LANCZOS4 3) the luma range shrinks by h, the chroma range by 2 h luma pixels.
            const double h = (double)g.halo, h2 = 2.0 * h;
            if (lo[0] >= 0.55 + h2 && hi[0] <= W - 1.55 - h2 && lo[1] >= 0.55 + h2 && hi[1] <= H - 1.55 - h2) flags |= kPieceInterior;
            // every luma and chroma tap outside in x, or in y
            if (hi[0] < -1.6 - h2 || lo[0] > W + 0.55 + h2 || hi[1] < -1.6 - h2 || lo[1] > H + 0.55 + h2) flags |= kPieceOutside;
            if (!(flags & kPieceOutside)) {
                // source rectangle of the taps (for the shared-memory staging): luma taps floor(m) - h ...
                // floor(m) + 1 + h; chroma coordinate = (mean - 0.5) / 2
                box.x0 = (int16_t)(floor(lo[0] - 0.01) - h); box.x1 = (int16_t)(floor(hi[0] + 0.01) + 1.0 + h);
                box.y0 = (int16_t)(floor(lo[1] - 0.01) - h); box.y1 = (int16_t)(floor(hi[1] + 0.01) + 1.0 + h);
                box.cx0 = (int16_t)(floor((lo[0] - 0.5) * 0.5 - 0.01) - h); box.cx1 = (int16_t)(floor((hi[0] - 0.5) * 0.5 + 0.01) + 1.0 + h);
                box.cy0 = (int16_t)(floor((lo[1] - 0.5) * 0.5 - 0.01) - h); box.cy1 = (int16_t)(floor((hi[1] - 0.5) * 0.5 + 0.01) + 1.0 + h);
            }
        }
        rec.flags = flags;
        rec.pad = 0;
        rec.box = box;
        PieceStage st = {0, 0, 0, 0, 0, 0, 0, 0};
        if ((flags & kPiecePoly) && !(flags & kPieceOutside)) {
            const int lx0 = box.x0 & ~15, wb = (box.x1 - lx0 + 16) & ~15;
            const int cbx0 = (2 * box.cx0) & ~15, cwb = (2 * box.cx1 + 2 - cbx0 + 15) & ~15;
            const int nrows = (box.y1 - box.y0 + 4) & ~3, cnrows = (box.cy1 - box.cy0 + 4) & ~3;  // rows, rounded up to whole 4-row boxes
            const int want = max(wb, cwb);
            const int pl128 = (want + 127) & ~127, pl32 = max(kStageMinPitch, (want + 31) & ~31);
            const int pl64 = (((want - 64 + 127) & ~127) + 64 < kStageMinPitch) ? kStageMinPitch + 64 : ((want - 64 + 127) & ~127) + 64;
            const int pl = (g.pitch64 && pl64 <= kStageMaxPitch && pl64 * (nrows + cnrows) <= g.tile_cap) ? pl64
                           : ((pl128 <= kStageMaxPitch && pl128 * (nrows + cnrows) <= g.tile_cap) ? pl128 : pl32);
            if (pl <= kStageMaxPitch && nrows > 0 && cnrows > 0 && nrows < 65536 && cnrows < 65536) {
                st.lx0 = (int16_t)lx0; st.by0 = box.y0; st.cbx0 = (int16_t)cbx0; st.cy0 = box.cy0;
                st.pl = (uint16_t)pl; st.nrows = (uint16_t)nrows; st.cnrows = (uint16_t)cnrows;
            }
        }
        rec.stage = st;
    }
    __syncthreads();
    if (tid < np && bad[tid]) {  // failed the accuracy certificate: per-pixel evaluation
        recs[tid].flags = 0;
        recs[tid].box = PieceBox{0, -1, 0, -1, 0, -1, 0, -1};
        recs[tid].stage = PieceStage{0, 0, 0, 0, 0, 0, 0, 0};
    }
    __syncthreads();
    // coalesced copy of the records (240 bytes each) to the table
    PieceRec* out = table + ((size_t)frame * npy + py) * npx + p0;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(recs);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(out);
    for (int i = tid; i < np * (int)(sizeof(PieceRec) / 4); i += kThreads) d32[i] = s32[i];
}

}  // namespace

cudaError_t launch_build_pieces(const GeomD& g, const PieceBasis& basis, const float* rots,
                                const float* rot0, int n_frames, PieceRec* table, cudaStream_t st)
{
    RotF r0{};
    if (rot0)
        for (int i = 0; i < 9; ++i) r0.r[i] = rot0[i];
    dim3 grid((pieces_x(g.out_w) + kChunk - 1) / kChunk, pieces_y(g.out_h, g.piece_h), n_frames);
    build_pieces_kernel<<<grid, kThreads, 0, st>>>(g, basis, rots, r0, table);
    return cudaGetLastError();
}

}  // namespace vaw
