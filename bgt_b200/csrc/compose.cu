// compose.cu -- composite plane-0 maps of row groups, for the QUERY phase of the split scan.
//
// A row's effect on a rank is a piecewise translation (one piece per run: rank' = rank + delta, pbwt.c:142-153).
// The composition of COMP_K consecutive rows is again a piecewise translation; every run boundary of every row
// has exactly one pre-image in the coordinates in front of the group (the maps are bijections), so the composite
// has at most 1 + (sum of the rows' runs) pieces.  A (column,row) query that only needs its rank at its own row can
// then cross a whole group with ONE binary search instead of COMP_K -- the bits of the rows in between are not
// needed by the count-only scan.
//
// One CTA per (checkpoint block, row group); groups are independent of each other and of the query, so the tables
// are built once per resident PBF (lazily, at the first scan that wants them) and cached with it.
// The same kernel also builds INVERSE composites (output coordinates -> input coordinates, rows composed in reverse)
// of the plane-1 view rows for plane1_select_kernel: P.inverse = 1, records of the view (rle at +9, n1 of plane 1).
//   out: comp_start[blk][g][cap] (piece starts, ascending, padded with 0xffffffff to a multiple of 4),
//        comp_delta[blk][g][COMP_CAP], comp_n[blk][g] = number of pieces (padded), 0 = not available (too many
//        pieces or runs for the staging buffers -> the walk falls back to row-by-row for that group).
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

constexpr int CP_NT = 256, CP_NW = CP_NT / 32;
constexpr int CP_RUNS = 4096;      // merged runs of all rows of a group held in shared memory

__device__ __forceinline__ uint32_t cp_rle_len(uint32_t c) { const uint32_t v = c >> 1; return (v & 15u) << ((v >> 4) << 2); }
__device__ __forceinline__ uint32_t cp_ld_u32_unaligned(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	const uint32_t lo = w[0];
	if (sh == 0) return lo;
	return __funnelshift_r(lo, w[1], sh);
}

__global__ void __launch_bounds__(CP_NT) pbwt_compose_kernel(const ComposeParams P)
{
	extern __shared__ __align__(16) uint8_t sm[];
	uint32_t *runs_s = (uint32_t*)sm;                      // [CP_RUNS] run starts of all rows (row after row)
	int32_t  *runs_d = (int32_t*)(runs_s + CP_RUNS);       // [CP_RUNS] run deltas
	uint32_t *A_s = (uint32_t*)(runs_d + CP_RUNS);         // [COMP_CAP + 1] piece starts (input coordinates)
	uint32_t *A_c = A_s + COMP_CAP + 1;                    // [COMP_CAP] piece positions in current coordinates
	uint32_t *B_s = A_c + COMP_CAP;                        // second list
	uint32_t *B_c = B_s + COMP_CAP + 1;
	__shared__ int row_beg[COMP_K + 1], row_nz[COMP_K];                    // first run of every row in runs_s (row_beg[j+1]-row_beg[j] = runs; 0 runs = identity)
	__shared__ int warp_tot[CP_NW];
	__shared__ int s_fail, s_n;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int BS = 1 << P.shift;
	const int n_grp = P.n_grp;
	const int blk = P.blk_list[blockIdx.x / n_grp], g = blockIdx.x % n_grp;
	const uint32_t m = (uint32_t)P.m;
	const long long rbase = P.row_base ? P.row_base[blk] : (long long)blk * BS;   // first row of the block in the per-row arrays
	const uint64_t *roff = P.rowoff + (P.row_base ? rbase + blk : (long long)blk * (BS + 1));
	const int r_lo = g * COMP_K;
	int nrow = P.rows_in_blk[blk] - r_lo;
	if (nrow > COMP_K) nrow = COMP_K;
	const size_t slot = ((size_t)blk * n_grp + g);
	const int cap = P.cap;
	if (nrow < COMP_K) { if (tid == 0) P.comp_n[slot] = 0; return; }   // partial last group: never crossed as a whole
	if (tid == 0) { s_fail = 0; s_n = 0; }
	__syncthreads();

	// ---- per-row merged run tables: warps take rows round-robin; run counts first (to place the rows), then the tables
	for (int pass = 0; pass < 2; ++pass) {
		for (int j = warp; j < nrow; j += CP_NW) {
			const uint8_t *rec = P.img + roff[r_lo + j];
			const uint32_t l = cp_ld_u32_unaligned(rec + P.rle_off - 4);
			const uint8_t *rle = rec + P.rle_off;
			const uint32_t n1 = P.n1[(size_t)(rbase + r_lo + j) * P.n1_step + P.n1_plane];
			const bool triv = (n1 == 0 || n1 == m);
			uint32_t tot = 0, ones = 0, nrun = 0, nzr = 0, prev_bit = 2;
			const int base_out = pass ? row_beg[j] : 0;
			const int nz_row = pass ? row_nz[j] : 0;
			if (!triv) {
				for (uint32_t base = 0; base < l; base += 32) {
					const uint32_t i = base + lane;
					const uint32_t c = i < l ? rle[i] : 0u;
					const uint32_t L = cp_rle_len(c), b = c & 1u, L1 = b ? L : 0u;
					uint32_t x = L, y = L1;
					#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const uint32_t tx = __shfl_up_sync(0xffffffffu, x, d), ty = __shfl_up_sync(0xffffffffu, y, d);
						if (lane >= d) { x += tx; y += ty; }
					}
					const uint32_t start = tot + x - L, ones_before = ones + y - L1;
					// a run starts at a byte of non-zero length whose bit differs from the previous non-empty byte
					const uint32_t valid = __ballot_sync(0xffffffffu, L > 0);
					const uint32_t bitm = __ballot_sync(0xffffffffu, b != 0);
					const uint32_t below = valid & ((1u << lane) - 1u);
					uint32_t pb = prev_bit;
					if (below) pb = (bitm >> (31 - __clz(below))) & 1u;
					const bool is_start = L > 0 && pb != b;
					const uint32_t sm_ = __ballot_sync(0xffffffffu, is_start);
					const uint32_t zm_ = sm_ & ~bitm;                      // run starts of 0-runs
					if (pass && is_start) {
						const int32_t delta = b ? (int32_t)((m - n1) - (start - ones_before)) : -(int32_t)ones_before;
						if (!P.inverse) {
							const int k = base_out + (int)nrun + __popc(sm_ & ((1u << lane) - 1u));
							if (k < CP_RUNS) { runs_s[k] = start; runs_d[k] = delta; }
						} else { // inverse map: runs ordered by where they land (0-runs, then 1-runs), translated back
							const uint32_t lt = (1u << lane) - 1u;
							const int k = b ? base_out + nz_row + (int)(nrun - nzr) + __popc(sm_ & bitm & lt) : base_out + (int)nzr + __popc(zm_ & lt);
							if (k < CP_RUNS) { runs_s[k] = start + (uint32_t)delta; runs_d[k] = -delta; }
						}
					}
					nrun += __popc(sm_);
					nzr += __popc(zm_);
					if (valid) prev_bit = (bitm >> (31 - __clz(valid))) & 1u;
					tot += __shfl_sync(0xffffffffu, x, 31);
					ones += __shfl_sync(0xffffffffu, y, 31);
				}
			}
			if (!pass && lane == 0) { row_beg[j + 1] = (int)nrun; row_nz[j] = (int)nzr; }   // counts for now
		}
		__syncthreads();
		if (!pass) {
			if (tid == 0) {
				int acc = 0;
				row_beg[0] = 0;
				for (int j = 0; j < nrow; ++j) { const int c = row_beg[j + 1]; acc += c; row_beg[j + 1] = acc; }
				if (acc > CP_RUNS) s_fail = 1;
			}
			__syncthreads();
			if (s_fail) { if (tid == 0) P.comp_n[slot] = 0; return; }
		}
	}

	// ---- compose row after row
	uint32_t *Ls = A_s, *Lc = A_c, *Ns = B_s, *Nc = B_c;
	int n = 1;
	if (tid == 0) { Ls[0] = 0; Lc[0] = 0; Ls[1] = m; }
	__syncthreads();
	for (int jj = 0; jj < nrow; ++jj) {
		const int j = P.inverse ? nrow - 1 - jj : jj;           // inverse composite: undo the last row first
		const int rb = row_beg[j], nr = row_beg[j + 1] - rb;
		if (nr == 0) continue;                                  // constant row: identity (pbwt.c:75-77)
		const uint32_t *rs = runs_s + rb;
		const int32_t *rd = runs_d + rb;
		// pieces per thread: contiguous chunk, so that the output stays in input order
		const int per = (n + CP_NT - 1) / CP_NT, p0 = tid * per, p1 = p0 + per < n ? p0 + per : n;
		int mine = 0;
		for (int p = p0; p < p1; ++p) {
			const uint32_t c = Lc[p], e = c + (Ls[p + 1] - Ls[p]);
			int lo = 0, hi = 0;                                  // lo = last run with start <= c ; hi = last run with start < e
			for (int len = nr; len > 1;) { const int half = len >> 1; lo += rs[lo + half] <= c ? half : 0; hi += rs[hi + half] < e ? half : 0; len -= half; }
			mine += hi - lo + 1;
		}
		int x = mine;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
		if (lane == 31) warp_tot[warp] = x;
		__syncthreads();
		int off = x - mine;
		for (int w = 0; w < warp; ++w) off += warp_tot[w];
		if (tid == CP_NT - 1) { s_n = off + mine; if (off + mine > cap - 4) s_fail = 1; }
		__syncthreads();
		if (s_fail) { if (tid == 0) P.comp_n[slot] = 0; return; }
		for (int p = p0; p < p1; ++p) {
			const uint32_t s0 = Ls[p], c = Lc[p], e = c + (Ls[p + 1] - s0);
			int lo = 0, hi = 0;
			for (int len = nr; len > 1;) { const int half = len >> 1; lo += rs[lo + half] <= c ? half : 0; hi += rs[hi + half] < e ? half : 0; len -= half; }
			Ns[off] = s0; Nc[off] = c + (uint32_t)rd[lo]; ++off;
			for (int k = lo + 1; k <= hi; ++k) { const uint32_t s = rs[k]; Ns[off] = s0 + (s - c); Nc[off] = s + (uint32_t)rd[k]; ++off; }
		}
		n = s_n;
		__syncthreads();
		if (tid == 0) Ns[n] = m;
		uint32_t *t;
		t = Ls; Ls = Ns; Ns = t;
		t = Lc; Lc = Nc; Nc = t;
		__syncthreads();
	}
	// ---- write out, padded to a multiple of 4 entries (16-byte TMA granularity)
	const int npad = (n + 3) & ~3;
	uint32_t *os = P.comp_start + slot * cap;
	int32_t *od = P.comp_delta + slot * cap;
	for (int p = tid; p < npad; p += CP_NT) {
		os[p] = p < n ? Ls[p] : 0xffffffffu;
		od[p] = p < n ? (int32_t)(Lc[p] - Ls[p]) : 0;
	}
	if (tid == 0) P.comp_n[slot] = npad;
}

size_t compose_smem_bytes() { return sizeof(uint32_t) * (2 * CP_RUNS + 4 * COMP_CAP + 2) + 64; }

cudaError_t launch_compose(const ComposeParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	const int n_grp = P.n_grp;
	const size_t smem = compose_smem_bytes();
	cudaError_t e = cudaFuncSetAttribute(pbwt_compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	pbwt_compose_kernel<<<(unsigned)((long long)n_blk * n_grp), CP_NT, smem, st>>>(P);
	return cudaGetLastError();
}

} // namespace b200
