/*
 * wn_kernels_format.cu -- the edge rows of edges-frames.xvg.dat formatted on the GPU (sm_100a).
 *
 * The reference writes one text row per edge through iostreams (src/waternetwork.cpp:638-652):
 *     setw(6) id1, setw(6) id2, then energy, length, cosine as fixed, precision 4, setw(8), '\n'
 * i.e. 37 bytes per row whenever the ids have at most 6 digits and the values fit 8 characters
 * (always the case for water: |energy| < 10, length <= 6.5, |cos| <= 1).  At ~16 MB of text per
 * 100k-atom frame this is what limits the reference once the search runs on the GPU (a host
 * formatter does ~0.4 GB/s per core), so the rows are produced here: one lane per edge, the
 * warp's 32 rows assembled in shared memory and copied out contiguously.  Rounding is the same
 * as glibc's printf: v*10^4 is exact in double for a float v, rint() rounds half to even.
 * A row that does not fit the fixed width raises the frame's flag; the caller then formats that
 * batch on the host (EdgeFileWriter), so the bytes are identical either way.
 */
#include "wn_internal.cuh"

namespace {

constexpr int ROW_BYTES = 37;

/* value as "%8.4f"; returns false if it needs more than 8 characters or is not finite */
__device__ __forceinline__ bool wn_put_fixed4(char* dst, float v)
{
    const double a = fabs((double)v);
    if (!(a < 1000.0)) return false;                       /* also catches NaN and inf */
    const unsigned int scaled = (unsigned int)__double2ll_rn(a * 10000.0);   /* exact product, half-to-even */
    const unsigned int ip = scaled / 10000u;
    unsigned int fp = scaled - ip * 10000u;
    const bool neg = signbit(v);
    const int nd = ip >= 100u ? 3 : (ip >= 10u ? 2 : 1);
    if (nd + (neg ? 1 : 0) + 5 > 8) return false;
    dst[7] = (char)('0' + fp % 10u); fp /= 10u;
    dst[6] = (char)('0' + fp % 10u); fp /= 10u;
    dst[5] = (char)('0' + fp % 10u); fp /= 10u;
    dst[4] = (char)('0' + fp);
    dst[3] = '.';
    unsigned int t = ip;
    int pos = 2;
    for (int k = 0; k < nd; ++k) { dst[pos--] = (char)('0' + t % 10u); t /= 10u; }
    if (neg) dst[pos--] = '-';
    while (pos >= 0) dst[pos--] = ' ';
    return true;
}

/* unsigned as "%6u"; false if it needs more than 6 digits */
__device__ __forceinline__ bool wn_put_uint6(char* dst, unsigned int v)
{
    if (v > 999999u) return false;
    int pos = 5;
    do { dst[pos--] = (char)('0' + v % 10u); v /= 10u; } while (v);
    while (pos >= 0) dst[pos--] = ' ';
    return true;
}

__global__ void __launch_bounds__(256)
wn_k_format_rows(const wn_edge* __restrict__ edges, const unsigned long long* __restrict__ frame_off,
                 const unsigned long long* __restrict__ row_base, const unsigned int* __restrict__ site_ids,
                 char* __restrict__ text, int* __restrict__ frame_flags)
{
    __shared__ char s_rows[8][32 * ROW_BYTES];
    const int f = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long e_begin = frame_off[f], e_end = frame_off[f + 1];
    const unsigned long long e0 = e_begin + ((unsigned long long)blockIdx.x * 8 + warp) * 32;
    if (e0 >= e_end) return;
    const int nrows = (int)min((unsigned long long)32, e_end - e0);
    char* row = s_rows[warp] + lane * ROW_BYTES;
    bool ok = true;
    if (lane < nrows) {
        const wn_edge e = edges[e0 + lane];
        const unsigned int id1 = site_ids ? site_ids[e.i] : (unsigned int)e.i;
        const unsigned int id2 = site_ids ? site_ids[e.j] : (unsigned int)e.j;
        ok = wn_put_uint6(row, id1) & wn_put_uint6(row + 6, id2) & wn_put_fixed4(row + 12, e.energy) &
             wn_put_fixed4(row + 20, e.length) & wn_put_fixed4(row + 28, e.cosine);
        row[36] = '\n';
    }
    if (__any_sync(0xffffffffu, !ok) && lane == 0) atomicOr(frame_flags + f, 1);
    __syncwarp();
    /* contiguous copy of the warp's rows */
    char* dst = text + row_base[f] + (e0 - e_begin) * ROW_BYTES;
    const int nbytes = nrows * ROW_BYTES;
    const char* src = s_rows[warp];
    /* bytes up to the first 4-byte boundary of dst, whole words, then the tail */
    const int head = min(nbytes, (int)((4 - ((size_t)dst & 3)) & 3));
    if (lane < head) dst[lane] = src[lane];
    const int nwords = (nbytes - head) >> 2;
    for (int w = lane; w < nwords; w += 32) {
        const char* s = src + head + 4 * w;
        const unsigned int word = (unsigned int)(unsigned char)s[0] | ((unsigned int)(unsigned char)s[1] << 8) |
                                  ((unsigned int)(unsigned char)s[2] << 16) | ((unsigned int)(unsigned char)s[3] << 24);
        *reinterpret_cast<unsigned int*>(dst + head + 4 * w) = word;
    }
    const int done = head + 4 * nwords;
    if (lane < nbytes - done) dst[done + lane] = src[done + lane];
}

}  // namespace

void wn_launch_format_rows(const wn_edge* edges, const unsigned long long* frame_off,
                           const unsigned long long* row_base, const unsigned int* site_ids, int n_frames,
                           unsigned long long max_edges_per_frame, char* text, int* frame_flags, cudaStream_t st)
{
    if (n_frames <= 0 || max_edges_per_frame == 0) return;
    dim3 g((unsigned)((max_edges_per_frame + 255) / 256), n_frames);
    wn_k_format_rows<<<g, 256, 0, st>>>(edges, frame_off, row_base, site_ids, text, frame_flags);
}
