// compose.cu -- composite plane-0 maps of row groups, for the QUERY phase of the split scan.
//
// A row's effect on a rank is a piecewise translation (one piece per run: rank' = rank + delta, pbwt.c:142-153).
// The composition of COMP_K consecutive rows is again a piecewise translation; every run boundary of every row
// has exactly one pre-image in the coordinates in front of the group (the maps are bijections), so the composite
// has at most 1 + (sum of the rows' runs) pieces.  A (column,row) query that only needs its rank at its own row can
// then cross a whole group with ONE binary search instead of COMP_K -- the bits of the rows in between are not
// needed by the count-only scan.
//
// The COMP_K row maps are composed pairwise in a tree (log2 COMP_K levels), each level one pass over all pieces.
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
constexpr int CP_RUNS_BIG = 4096, CP_RUNS_SMALL = 2752;   // merged runs of all rows of a group held in shared memory: the small
// variant (49 KB, 4 CTAs per SM) takes every group first, the big one (66 KB, 3 per SM) re-does the few that did not fit
constexpr int CP_CACHE = 10;       // look-ups per thread kept in registers between the two passes of a level (total/2/CP_NT ~ 8)

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

// Pieces of the maps of one level live back to back in (S, D): S = start of the piece in the map's input coordinates,
// D = translation.  off[i] .. off[i+1] are the pieces of map i.
template<int CP_RUNS>
__global__ void __launch_bounds__(CP_NT, CP_RUNS == CP_RUNS_SMALL ? 5 : 3) pbwt_compose_kernel(const ComposeParams P)
{
	extern __shared__ __align__(16) uint8_t sm[];
	constexpr int CPX = CP_RUNS + COMP_K + 8;
	uint32_t *S0 = (uint32_t*)sm;          int32_t *D0 = (int32_t*)(S0 + CPX);
	uint32_t *S1 = (uint32_t*)(D0 + CPX);  int32_t *D1 = (int32_t*)(S1 + CPX);
	__shared__ int offA[COMP_K + 1], offB[COMP_K + 1], row_nz[COMP_K];
	__shared__ int warp_tot[CP_NW];
	__shared__ int s_fail, s_n;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int BS = 1 << P.shift;
	const int n_grp = P.n_grp;
	const int blk = P.blk_list ? P.blk_list[blockIdx.x / n_grp] : P.blk_first + (int)(blockIdx.x / n_grp), g = blockIdx.x % n_grp;
	const uint32_t m = (uint32_t)P.m;
	const long long rbase = P.row_base ? P.row_base[blk] : (long long)blk * BS;   // first row of the block in the per-row arrays
	const uint64_t *roff = P.rowoff + (P.row_base ? rbase + blk : (long long)blk * (BS + 1));
	const int r_lo = g * COMP_K;
	const int nrow = P.rows_in_blk[blk] - r_lo;
	const size_t slot = ((size_t)blk * n_grp + g);
	const int cap = P.cap;
	const bool inverse = P.inverse || g >= comp_first_inverse(blk, P.n_blk_res, P.rows_in_blk[blk], BS, n_grp, P.two_sided);
	if (CP_RUNS == CP_RUNS_BIG && P.retry && P.comp_n[slot] != 0) return;   // second launch: only the groups the small variant gave up on
	if (nrow < COMP_K || (P.blk_ok && !P.blk_ok[blk])) { if (tid == 0) P.comp_n[slot] = 0; return; }   // partial last group: never crossed as a whole
	if (tid == 0) { s_fail = 0; s_n = 0; }
	__syncthreads();

	// The records of the group lie back to back in the image: stage them once in the second buffer (idle until level 1)
	// so both passes of level 0 read shared memory.
	const uint64_t g_beg = roff[r_lo] & ~(uint64_t)3, g_end = roff[r_lo + COMP_K];
	const bool staged = g_end - g_beg + 8 <= (uint64_t)CPX * 8;
	if (staged) {
		const uint32_t nw = (uint32_t)((g_end - g_beg + 3) >> 2) + 1;        // one word of slack for the funnel-shift read
		const uint32_t *src = (const uint32_t*)(P.img + g_beg);
		for (uint32_t i = tid; i < nw; i += CP_NT) S1[i] = src[i];           // the images end with 64 bytes of padding
		__syncthreads();
	}

	// ---- level 0: one map per row, in the order they are applied (inverse composite: last row first).  Warps take
	// rows round-robin; run counts first (to place the maps), then the tables.  A constant row is the identity.
	auto place_maps = [&]() { // run counts in offA[1..] -> offsets; too many for the buffers: give up (the caller retries / walks rows)
		if (tid == 0) {
			int acc = 0;
			offA[0] = 0;
			for (int k = 0; k < COMP_K; ++k) { const int c = offA[k + 1]; acc += c; offA[k + 1] = acc; }
			if (acc > CP_RUNS + COMP_K) s_fail = 1;
		}
		__syncthreads();
	};
	const bool have_counts = P.nrun != nullptr;   // counted while loading (rowmeta_kernel): no counting pass
	if (have_counts) {
		if (tid < COMP_K) {
			const size_t ri = (size_t)(rbase + r_lo + tid);
			const uint32_t n1 = P.n1[ri * P.n1_step + P.n1_plane], nr = P.nrun[ri] & 0x7fffffffu, first = P.nrun[ri] >> 31;
			const int k_map = inverse ? COMP_K - 1 - tid : tid;
			offA[k_map + 1] = (n1 == 0 || n1 == m) ? 1 : (int)nr;
			row_nz[k_map] = (int)(first ? nr >> 1 : (nr + 1) >> 1);         // runs alternate: the 0-runs among them
		}
		__syncthreads();
		place_maps();
		if (s_fail) { if (tid == 0) P.comp_n[slot] = 0; return; }
	}
	for (int pass = have_counts ? 1 : 0; pass < 2; ++pass) {
		for (int j = warp; j < COMP_K; j += CP_NW) {
			const int k_map = inverse ? COMP_K - 1 - j : j;
			const uint8_t *rec = staged ? (const uint8_t*)S1 + (roff[r_lo + j] - g_beg) : P.img + roff[r_lo + j];
			const uint32_t l = cp_ld_u32_unaligned(rec + P.rle_off - 4);
			const uint8_t *rle = rec + P.rle_off;
			const uint32_t n1 = P.n1[(size_t)(rbase + r_lo + j) * P.n1_step + P.n1_plane];
			const bool triv = (n1 == 0 || n1 == m);
			uint32_t tot = 0, ones = 0, nrun = 0, nzr = 0, prev_bit = 2;
			const int base_out = pass ? offA[k_map] : 0;
			const int nz_row = pass ? row_nz[k_map] : 0;
			if (triv) {
				if (pass && lane == 0) { S0[base_out] = 0; D0[base_out] = 0; }
				nrun = 1;
			} else {
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
					const uint32_t lt = (1u << lane) - 1u;
					const uint32_t below = valid & lt;
					uint32_t pb = prev_bit;
					if (below) pb = (bitm >> (31 - __clz(below))) & 1u;
					const bool is_start = L > 0 && pb != b;
					const uint32_t sm_ = __ballot_sync(0xffffffffu, is_start);
					const uint32_t zm_ = sm_ & ~bitm;                      // run starts of 0-runs
					if (pass && is_start) {
						const int32_t delta = b ? (int32_t)((m - n1) - (start - ones_before)) : -(int32_t)ones_before;
						if (!inverse) {
							const int k = base_out + (int)nrun + __popc(sm_ & lt);
							S0[k] = start; D0[k] = delta;
						} else { // inverse map: runs ordered by where they land (0-runs, then 1-runs), translated back
							const int k = b ? base_out + nz_row + (int)(nrun - nzr) + __popc(sm_ & bitm & lt) : base_out + (int)nzr + __popc(zm_ & lt);
							S0[k] = start + (uint32_t)delta; D0[k] = -delta;
						}
					}
					nrun += __popc(sm_);
					nzr += __popc(zm_);
					if (valid) prev_bit = (bitm >> (31 - __clz(valid))) & 1u;
					tot += __shfl_sync(0xffffffffu, x, 31);
					ones += __shfl_sync(0xffffffffu, y, 31);
				}
			}
			if (!pass && lane == 0) { offA[k_map + 1] = (int)nrun; row_nz[k_map] = (int)nzr; }   // counts for now
		}
		__syncthreads();
		if (!pass) {
			place_maps();
			if (s_fail) { if (tid == 0) P.comp_n[slot] = 0; return; }
		}
	}

	// ---- compose pairwise, level by level: result i = (map 2i+1) after (map 2i).  Every piece of map 2i is cut at the
	// starts of map 2i+1 that fall inside its image; one scan over all pieces of the level places the results.
	// Work items are the pieces of the FIRST maps of the pairs only (those of the second maps are just looked up), dealt
	// out evenly; the look-up of a piece (first cut + number of cuts) is kept in registers between the counting and the
	// placing pass.
	uint32_t *Sa = S0, *Sb = S1;
	int32_t *Da = D0, *Db = D1;
	int *off = offA, *noff = offB;
	for (int nmaps = COMP_K; nmaps > 1; nmaps >>= 1) {
		int total_even = 0;
		for (int j = 0; j < nmaps; j += 2) total_even += off[j + 1] - off[j];
		const int per = (total_even + CP_NT - 1) / CP_NT;
		const int w0 = tid * per, w1 = w0 + per < total_even ? w0 + per : total_even;
		// first item: pair j0 and piece p_first
		int j0 = 0, p_first = 0;
		{
			int acc = 0;
			for (; j0 < nmaps; j0 += 2) { const int c = off[j0 + 1] - off[j0]; if (w0 < acc + c) break; acc += c; }
			p_first = j0 < nmaps ? off[j0] + (w0 - acc) : 0;
		}
		// piece p of map o against map o+1: lo = last start <= image start, cnt = pieces of the result
		auto cut = [&](int p, int o, int &lo, int &cnt) {
			const uint32_t s0 = Sa[p], cur = s0 + (uint32_t)Da[p];
			const uint32_t e = cur + ((p + 1 < off[o + 1] ? Sa[p + 1] : m) - s0);
			const uint32_t *rs = Sa + off[o + 1];
			const int nr = off[o + 2] - off[o + 1];
			// l = last start <= cur, h = last start < e: two searches over the same starts in lock step (the cuts of a piece are
			// few but unevenly spread, so counting them one by one keeps a warp waiting for its unluckiest lane)
			int l = 0, h = 0;
			for (int len = nr; len > 1;) {
				const int half = len >> 1;
				const uint32_t vl = rs[l + half], vh = rs[h + half];
				l += vl <= cur ? half : 0;
				h += vh < e ? half : 0;
				len -= half;
			}
			lo = l; cnt = h - l + 1;
		};
		int lo_c[CP_CACHE], cnt_c[CP_CACHE];
		int mine = 0;
		{
			int o = j0, p = p_first;
			#pragma unroll
			for (int q = 0; q < CP_CACHE; ++q) {
				lo_c[q] = 0; cnt_c[q] = 0;
				if (w0 + q < w1) {
					cut(p, o, lo_c[q], cnt_c[q]);
					mine += cnt_c[q];
					if (++p == off[o + 1]) { o += 2; p = off[o < nmaps ? o : 0]; }
				}
			}
			for (int w = w0 + CP_CACHE; w < w1; ++w) {
				int l, c;
				cut(p, o, l, c);
				mine += c;
				if (++p == off[o + 1]) { o += 2; p = off[o < nmaps ? o : 0]; }
			}
		}
		int x = mine;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
		if (lane == 31) warp_tot[warp] = x;
		__syncthreads();
		int out = x - mine, n_level = 0;                            // every thread adds up the warp totals itself: no second barrier
		#pragma unroll
		for (int w = 0; w < CP_NW; ++w) { const int t = warp_tot[w]; n_level += t; out += w < warp ? t : 0; }
		if (n_level > CPX - 4 || (nmaps == 2 && n_level > cap - 4)) { if (tid == 0) P.comp_n[slot] = 0; return; }   // (uniform)
		{
			int o = j0, p = p_first;
			auto place = [&](int lo, int cnt) {
				if (p == off[o]) noff[o >> 1] = out;                   // first piece of a pair's first map = start of the result map
				const uint32_t s0 = Sa[p], cur = s0 + (uint32_t)Da[p];
				const uint32_t *rs = Sa + off[o + 1];
				const int32_t *rd = Da + off[o + 1];
				Sb[out] = s0; Db[out] = (int32_t)(cur + (uint32_t)rd[lo] - s0); ++out;
				for (int k = lo + 1; k < lo + cnt; ++k) {
					const uint32_t st = rs[k], in_s = s0 + (st - cur);
					Sb[out] = in_s; Db[out] = (int32_t)(st + (uint32_t)rd[k] - in_s); ++out;
				}
				if (++p == off[o + 1]) { o += 2; p = off[o < nmaps ? o : 0]; }
			};
			#pragma unroll
			for (int q = 0; q < CP_CACHE; ++q)
				if (w0 + q < w1) place(lo_c[q], cnt_c[q]);
			for (int w = w0 + CP_CACHE; w < w1; ++w) {
				int l, c;
				cut(p, o, l, c);
				place(l, c);
			}
		}
		if (tid == 0) noff[nmaps >> 1] = n_level;
		__syncthreads();                                             // (also: warp_tot is free for the next level)
		{ uint32_t *t = Sa; Sa = Sb; Sb = t; }
		{ int32_t *t = Da; Da = Db; Db = t; }
		{ int *t = off; off = noff; noff = t; }
	}
	// ---- write out, padded to a multiple of 4 entries (16-byte TMA granularity)
	const int n = off[1];
	const int npad = (n + 3) & ~3;
	uint32_t *os = P.comp_start + slot * cap;
	int32_t *od = P.comp_delta + slot * cap;
	for (int p = tid; p < npad; p += CP_NT) {
		os[p] = p < n ? Sa[p] : 0xffffffffu;
		od[p] = p < n ? Da[p] : 0;
	}
	// bucket directory for the walk's look-up: dir[b] = last piece that starts at or below rank b << dir_shift, so the piece
	// of a rank r lies in [dir[r >> s], dir[(r >> s) + 1]] -- a window of a few entries instead of a search over all n
	if (P.comp_dir) {
		uint16_t *dir = P.comp_dir + slot * COMP_DIR_STRIDE;
		const int sh = P.dir_shift;
		const uint32_t nb = ((m - 1) >> sh) + 1;                         // buckets that hold a rank
		for (int p = tid; p < n; p += CP_NT) {
			const uint32_t s0 = Sa[p], s1 = p + 1 < n ? Sa[p + 1] : m;
			const uint32_t b0 = (s0 + (1u << sh) - 1u) >> sh, b1 = (s1 + (1u << sh) - 1u) >> sh;   // buckets whose base lies in [s0, s1)
			for (uint32_t b = b0; b < b1 && b < nb; ++b) dir[b] = (uint16_t)p;
		}
		for (int b = (int)nb + tid; b < P.dir_n; b += CP_NT) dir[b] = (uint16_t)(n - 1);    // sentinel (+ padding)
	}
	if (tid == 0) P.comp_n[slot] = npad;
}

size_t compose_smem_bytes() { return sizeof(uint32_t) * 4 * (CP_RUNS_BIG + COMP_K + 8) + 64; }
static size_t compose_smem_of(int runs) { return sizeof(uint32_t) * 4 * (size_t)(runs + COMP_K + 8) + 64; }

cudaError_t launch_compose(const ComposeParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	const int n_grp = P.n_grp;
	const unsigned grid = (unsigned)((long long)n_blk * n_grp);
	cudaError_t e = cudaFuncSetAttribute(pbwt_compose_kernel<CP_RUNS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)compose_smem_of(CP_RUNS_SMALL));
	if (e == cudaSuccess) e = cudaFuncSetAttribute(pbwt_compose_kernel<CP_RUNS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)compose_smem_of(CP_RUNS_BIG));
	if (e != cudaSuccess) return e;
	ComposeParams Q = P;
	Q.retry = 0;
	pbwt_compose_kernel<CP_RUNS_SMALL><<<grid, CP_NT, compose_smem_of(CP_RUNS_SMALL), st>>>(Q);
	Q.retry = 1;
	pbwt_compose_kernel<CP_RUNS_BIG><<<grid, CP_NT, compose_smem_of(CP_RUNS_BIG), st>>>(Q);
	return cudaGetLastError();
}

} // namespace b200
