// marginal.cu -- per-group plane-0 marginals for the split scan of grouped full-cohort queries (`-s A -s B -f ... -G`).
//
// With several sample groups the number of ALT codes per group is (ones of the plane-0 row inside the group) minus
// (other-ALT codes inside the group).  The second term comes from the plane-1 queries (plane1.cu + WALK_QUERY); the
// first one needs no per-column state either: keep, per checkpoint block and group, ONE BIT per rank -- "the column
// at this rank belongs to the group" -- and push that bit vector through the same stable partition the PBWT applies
// to the permutation (pbwt.c:79-88): the bits under the row's 0-runs move to the front, the bits under its 1-runs
// behind them, both in order.  The popcount of the part that moved behind is the group's number of ones.
// m/32 words per row instead of m rank updates.
//
// A CTA owns one bit vector (two buffers in shared memory); the next vector is GATHERED through the row's inverse run
// table (sorted by landing position).  Rows are a dependent chain, so a block is cut into segments of 8 row groups: the
// seed kernel pushes the block's start vector through the composite maps of the row groups (compose.cu; one scatter
// per 32 rows) and stores it in front of every segment, then grid = (blocks x segments, groups-1) CTAs of 256 threads walk
// their 256 rows side by side.  Without resident composite maps: grid = (blocks, groups-1), one CTA of 1024 threads per
// block.  Rows are staged in tiles like in the walk kernel (plain cooperative loads).
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

constexpr int MG_NT = 1024, MG_RAW = 4096, MG_TMAX = 32;      // one CTA per block
constexpr int MG_NT_SEG = 256, MG_RAW_SEG = 2048;              // one CTA per segment of a block

__device__ __forceinline__ uint32_t mg_rle_len(uint32_t c) { const uint32_t v = c >> 1; return (v & 15u) << ((v >> 4) << 2); }

__device__ __forceinline__ uint32_t mg_ld_u32_unaligned(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	const uint32_t lo = w[0];
	if (sh == 0) return lo;
	return __funnelshift_r(lo, w[1], sh);
}

// the block's start vector: bit i = group of the column at rank i under the plane-0 snapshot (pbwt.c:298-300)
template<int NT>
__device__ __forceinline__ void mg_snapshot_vector(const MarginalParams &P, int blk, int g, int wpad, uint32_t *V0, uint32_t *V1)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t m = (uint32_t)P.m;
	const uint8_t *S0 = P.img + P.blkoff[blk] + 1;
	for (int w = warp; w < wpad; w += NT / 32) {
		const uint32_t i = (uint32_t)w * 32 + lane;
		bool in = false;
		if (i < m) { const uint32_t col = mg_ld_u32_unaligned(S0 + 4 * (size_t)i); in = col < m && P.tgrp[col] == g; }
		const uint32_t bits = __ballot_sync(0xffffffffu, in);
		if (lane == 0) { V0[w] = bits; V1[w] = 0; }
	}
}

// Rows [seg * seg_rows, (seg+1) * seg_rows) of one block and one group: NT threads, RAW bytes of staged records.  With
// one segment per block the vector starts at the snapshot; otherwise at the segment vector left by the seed kernel.
template<int NT, int RAW>
__global__ void __launch_bounds__(NT) pbwt_marginal_kernel(const MarginalParams P)
{
	constexpr int MG_NT = NT, MG_NW = NT / 32, MG_RAW = RAW;
	extern __shared__ __align__(16) uint8_t sm[];
	const int words = (P.m + 31) / 32, wpad = (words + 4 + 3) & ~3;   // (+ the word a funnel shift reads past the end)
	uint32_t *V0 = (uint32_t*)sm, *V1 = V0 + wpad;
	uint32_t *ts = V1 + wpad;                       // [MG_RAW] run starts (per RLE byte)
	int32_t *td = (int32_t*)(ts + MG_RAW);          // [MG_RAW] rank shift of the run
	uint8_t *raw = (uint8_t*)(td + MG_RAW);         // [MG_RAW + 16]
	uint32_t *r_off = (uint32_t*)(raw + MG_RAW + 16); // [MG_TMAX] offset of the plane-0 RLE of the tile's rows in raw
	uint32_t *r_len = r_off + MG_TMAX, *r_n1 = r_len + MG_TMAX;   // entries of the row's table; ones of the row
	int32_t *r_cnt = (int32_t*)(r_n1 + MG_TMAX);    // [MG_TMAX] ones of the group per row
	uint32_t *r_nz = (uint32_t*)(r_cnt + MG_TMAX);  // [MG_TMAX] entries that are 0-runs (they come first)
	uint32_t *r_nch = r_nz + MG_TMAX;               // [MG_TMAX] 32-word chunks of whole words over all entries
	uint16_t *cst = (uint16_t*)(r_nch + MG_TMAX);   // [MG_RAW] per entry: chunks in front of it
	__shared__ int s_nr;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int bi = (int)blockIdx.x / P.n_seg, seg = (int)blockIdx.x % P.n_seg;
	const int blk = P.blk_list ? P.blk_list[bi] : P.blk_first + bi, g = blockIdx.y;
	if (P.blk_ok && !P.blk_ok[blk]) return;
	if (P.blk_only && !P.blk_only[(size_t)bi * P.n_vec + g]) return;
	const int BS = 1 << P.shift;
	const uint32_t m = (uint32_t)P.m;
	const uint64_t *roff = P.rowoff + (size_t)blk * (BS + 1);
	const long long blk_row = P.blk_row0 + ((long long)blk << P.shift);
	int rows = P.rows_in_blk[blk];
	if (blk_row + rows > P.row_hi) rows = (int)(P.row_hi - blk_row);
	int row_a = 0;
	if (P.n_seg > 1) {
		const int seg_rows = P.seg_groups * COMP_K;
		if (P.seg_ok[(size_t)bi * P.n_vec + g]) { row_a = seg * seg_rows; if (row_a + seg_rows < rows) rows = row_a + seg_rows; }
		else if (seg > 0) return;                        // no segment vectors for this block: segment 0 takes all its rows
		if (row_a >= rows || blk_row + rows <= P.row_lo) return;
	}

	// ---- the start vector
	if (row_a == 0) mg_snapshot_vector<NT>(P, blk, g, wpad, V0, V1);
	else {
		if (P.vrec) {   // (the vectors of margpiece.cu: look-up records {64 bits, members in front}, one in front of every 32-row group)
			const int W64 = (int)marginal_rec_words64_dev(P.m);
			const uint4 *rec = P.vrec + (((size_t)bi * P.n_vec + g) * P.seg_slots + (size_t)seg * P.seg_slot_step) * (size_t)W64;
			for (int w = tid; w < wpad; w += MG_NT) {
				uint32_t v = 0;
				if ((w >> 1) < W64) { const uint4 r = rec[w >> 1]; v = (w & 1) ? r.y : r.x; }
				V0[w] = v; V1[w] = 0;
			}
		} else {
			const uint32_t *src = P.vseg + (((size_t)bi * P.n_vec + g) * P.seg_slots + (size_t)seg * P.seg_slot_step) * (size_t)wpad;
			for (int w = tid; w < wpad; w += MG_NT) { V0[w] = src[w]; V1[w] = 0; }
		}
	}
	__syncthreads();
	int total = 0;                                   // columns of the group (for all-ones rows)
	for (int w = tid; w < words; w += MG_NT) total += __popc(V0[w]);
	#pragma unroll
	for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
	if (tid < MG_TMAX) r_cnt[tid] = 0;
	__shared__ int s_total;
	if (tid == 0) s_total = 0;
	__syncthreads();
	if (lane == 0 && total) atomicAdd(&s_total, total);
	__syncthreads();
	const int group_cols = s_total;

	uint32_t *Vold = V0, *Vnew = V1;                 // Vnew is all zero here
	for (int r0 = row_a; r0 < rows;) {
		// ---- tile: as many rows as fit MG_RAW bytes (a single larger row is staged piecewise below)
		const uint64_t t_beg = roff[r0];
		if (warp == 0) { // (MG_TMAX == 32: one lane per candidate row)
			const bool fits = r0 + lane < rows && roff[r0 + lane + 1] - t_beg <= (uint64_t)MG_RAW;
			const uint32_t no = ~__ballot_sync(0xffffffffu, fits);
			const int k = no ? __ffs(no) - 1 : 32;
			if (lane == 0) s_nr = k > 0 ? k : 1;
		}
		__syncthreads();
		const int nr = s_nr;
		const bool big = roff[r0 + 1] - t_beg > (uint64_t)MG_RAW;
		if (!big) {
			const uint32_t nbytes = (uint32_t)(roff[r0 + nr] - t_beg);
			uint32_t my_n1 = 0, my_o = 0;                   // (issued with the staging loads: one round trip, not two)
			if (tid < nr) { my_n1 = P.n1[((size_t)blk * BS + r0 + tid) * 2]; my_o = (uint32_t)(roff[r0 + tid] - t_beg); }
			for (uint32_t i = tid; i < nbytes; i += MG_NT) raw[i] = P.img[t_beg + i];
			__syncthreads();
			if (tid < nr) {
				const uint32_t o = my_o;
				r_len[tid] = (uint32_t)raw[o + 1] | (uint32_t)raw[o + 2] << 8 | (uint32_t)raw[o + 3] << 16 | (uint32_t)raw[o + 4] << 24;
				r_off[tid] = o + 5;
				r_n1[tid] = my_n1;
			}
			__syncthreads();
			// per row the INVERSE run table of plane 0, one entry per run (RLE bytes of the same symbol merged): where the run's
			// bits LAND (0-runs in front in order, 1-runs behind the zt zeros in order, pbwt.c:79-88) and how far back their
			// source lies -- sorted by landing position, so the next vector can be gathered word by word
			for (int r = warp; r < nr; r += MG_NW) {
				const uint32_t n1 = r_n1[r], len = r_len[r], off = r_off[r];
				if (n1 == 0 || n1 == m) continue;
				const uint32_t zt = m - n1;
				const uint32_t lt = (1u << lane) - 1u;
				uint32_t nzr = 0, nrun = 0, prev_bit = 2;        // 0-runs, runs
				for (uint32_t base = 0; base < len; base += 32) {
					const uint32_t i = base + lane;
					const uint32_t c = i < len ? raw[off + i] : 0u;
					const uint32_t L = mg_rle_len(c), b = c & 1u;
					const uint32_t valid = __ballot_sync(0xffffffffu, L > 0), bitm = __ballot_sync(0xffffffffu, b != 0);
					const uint32_t below = valid & lt;
					const uint32_t pb = below ? (bitm >> (31 - __clz(below))) & 1u : prev_bit;
					const uint32_t sm_ = __ballot_sync(0xffffffffu, L > 0 && pb != b);
					nrun += __popc(sm_); nzr += __popc(sm_ & ~bitm);
					if (valid) prev_bit = (bitm >> (31 - __clz(valid))) & 1u;
				}
				uint32_t tot = 0, ones = 0, kz = 0, ko = 0;
				prev_bit = 2;
				for (uint32_t base = 0; base < len; base += 32) {
					const uint32_t i = base + lane;
					const uint32_t c = i < len ? raw[off + i] : 0u;
					const uint32_t L = mg_rle_len(c), b = c & 1u, L1 = b ? L : 0u;
					uint32_t x = L, y = L1;
					#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const uint32_t tx = __shfl_up_sync(0xffffffffu, x, d), ty = __shfl_up_sync(0xffffffffu, y, d);
						if (lane >= d) { x += tx; y += ty; }
					}
					const uint32_t start = tot + x - L, ones_before = ones + y - L1;
					const uint32_t valid = __ballot_sync(0xffffffffu, L > 0), bitm = __ballot_sync(0xffffffffu, b != 0);
					const uint32_t below = valid & lt;
					const uint32_t pb = below ? (bitm >> (31 - __clz(below))) & 1u : prev_bit;
					const bool is_start = L > 0 && pb != b;
					const uint32_t sm_ = __ballot_sync(0xffffffffu, is_start);
					const uint32_t zmask = sm_ & ~bitm, omask = sm_ & bitm;
					if (is_start) {
						const uint32_t dst = b ? zt + ones_before : start - ones_before;
						const uint32_t k = b ? nzr + ko + __popc(omask & lt) : kz + __popc(zmask & lt);
						ts[off + k] = dst; td[off + k] = (int32_t)(start - dst);
					}
					kz += __popc(zmask); ko += __popc(omask);
					if (valid) prev_bit = (bitm >> (31 - __clz(valid))) & 1u;
					tot += __shfl_sync(0xffffffffu, x, 31);
					ones += __shfl_sync(0xffffffffu, y, 31);
				}
				__syncwarp();
				// whole words of every entry, in chunks of 32 (one warp step each), as a running count in front of the entry
				uint32_t carry = 0;
				for (uint32_t base = 0; base < nrun; base += 32) {
					const uint32_t k = base + lane;
					uint32_t ch = 0;
					if (k < nrun) {
						const uint32_t nxt = k + 1 < nrun ? ts[off + k + 1] : m;
						const uint32_t wa = (ts[off + k] + 31u) >> 5, we = nxt >> 5;
						ch = we > wa ? (we - wa + 31u) >> 5 : 0u;
					}
					uint32_t x = ch;
					#pragma unroll
					for (int d = 1; d < 32; d <<= 1) { const uint32_t tx = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += tx; }
					if (k < nrun) cst[off + k] = (uint16_t)(carry + x - ch);
					carry += __shfl_sync(0xffffffffu, x, 31);
				}
				if (lane == 0) { r_len[r] = nrun; r_nz[r] = nzr; r_nch[r] = carry; }
			}
			__syncthreads();
		}
		for (int r = 0; r < nr; ++r) {
			uint32_t n1, n;
			const uint32_t *rts; const int32_t *rtd;
			if (!big) { n1 = r_n1[r]; n = r_len[r]; rts = ts + r_off[r]; rtd = td + r_off[r]; }
			else { n1 = P.n1[((size_t)blk * BS + r0) * 2]; n = 0; rts = ts; rtd = td; }
			int cnt = 0;
			if (n1 == 0) { /* nothing moves, no ones */ }
			else if (n1 == m) { if (tid == 0) r_cnt[r] = group_cols; }
			else if (!big) {
				// gather, in two passes.  (1) the words that lie inside ONE entry -- nearly all of them -- are a single funnel shift
				// of two source words: they are dealt to the warps in chunks of 32 consecutive words of one entry (warp-uniform
				// entry cursor, one word per lane).  (2) the words that contain an entry boundary (one thread per boundary) are
				// assembled piece by piece.  Every word is written, none needs clearing.
				const uint32_t zt = m - n1;
				auto word_of = [&](uint32_t w, uint32_t k) -> uint32_t {   // general case; k = last entry starting at or below 32w
					const uint32_t pos = w * 32, lim = pos + 32 < m ? pos + 32 : m;
					uint32_t word = 0, cur = pos;
					while (cur < lim) {
						const uint32_t next = k + 1 < n ? rts[k + 1] : m;
						const uint32_t stop = next < lim ? next : lim;
						if (stop > cur) {
							const uint32_t src = cur + (uint32_t)rtd[k], sw = src >> 5, take = stop - cur;
							const uint32_t v = __funnelshift_r(Vold[sw], Vold[sw + 1], src & 31u) & (take == 32 ? 0xffffffffu : ((1u << take) - 1u));
							word |= v << (cur - pos);
							cur = stop;
						}
						if (cur >= next) ++k;
					}
					return word;
				};
				{
					const uint16_t *rc = cst + r_off[r];
					const uint32_t T = r_nch[r], nz = r_nz[r];
					uint32_t j = T * (uint32_t)warp / MG_NW;
					const uint32_t j_hi = T * (uint32_t)(warp + 1) / MG_NW;       // this warp's share of the chunks
					if (j < j_hi) {
						uint32_t k = 0;                                            // entry of chunk j: the last one with rc[k] <= j
						for (uint32_t len = n; len > 1;) { const uint32_t half = len >> 1; k += rc[k + half] <= j ? half : 0u; len -= half; }
						while (j < j_hi) {
							const uint32_t nxt = k + 1 < n ? rts[k + 1] : m;
							const uint32_t wa = (rts[k] + 31u) >> 5, we = nxt >> 5;
							if (we <= wa) { ++k; continue; }                        // no whole word inside this entry
							const uint32_t base = rc[k];
							uint32_t c_end = (we - wa + 31u) >> 5;
							if (base + c_end > j_hi) c_end = j_hi - base;
							const uint32_t d = (uint32_t)rtd[k], sh = d & 31u;
							const bool ones = k >= nz;
							for (uint32_t c = j - base; c < c_end; ++c) {
								const uint32_t w = wa + c * 32 + lane;
								if (w < we) {
									const uint32_t sw = (w * 32 + d) >> 5;
									const uint32_t word = __funnelshift_r(Vold[sw], Vold[sw + 1], sh);
									Vnew[w] = word;
									if (ones) cnt += __popc(word);
								}
							}
							j = base + c_end;
							++k;
						}
					}
				}
				// second pass (no barrier: it reads the old vector and writes only words the first pass never touches): exactly one
				// thread per word that holds an entry boundary -- the one whose boundary is the first in it
				for (uint32_t i = tid; i + 1 < n; i += MG_NT) {
					const uint32_t p = rts[i + 1];
					if (p >= m || (p & 31u) == 0) continue;           // (a boundary at a word start leaves both words whole)
					const uint32_t w = p >> 5, pos = w * 32;
					if (rts[i] > pos) continue;                       // an earlier boundary lies in the same word
					uint32_t k = 0;
					for (uint32_t len = n; len > 1;) { const uint32_t half = len >> 1; k += rts[k + half] <= pos ? half : 0u; len -= half; }
					const uint32_t word = word_of(w, k);
					Vnew[w] = word;
					if (pos >= zt) cnt += __popc(word);
					else if (pos + 32 > zt) cnt += __popc(word >> (zt - pos));
				}
				if (tid == 0 && (m & 31u)) {                          // the ragged last word, if no boundary lies inside it
					const uint32_t w = (uint32_t)words - 1, pos = w * 32;
					uint32_t k = 0;
					for (uint32_t len = n; len > 1;) { const uint32_t half = len >> 1; k += rts[k + half] <= pos ? half : 0u; len -= half; }
					bool inside = false;
					for (uint32_t j = k + 1; j < n && rts[j] < m; ++j) if (rts[j] > pos) { inside = true; break; }
					if (!inside) {
						const uint32_t word = word_of(w, k);
						Vnew[w] = word;
						if (pos >= zt) cnt += __popc(word);
						else if (pos + 32 > zt) cnt += __popc(word >> (zt - pos));
					}
				}
			} else {
				// a row larger than the staging buffer: stream its plane-0 RLE in pieces; every piece covers a rank range
				for (int w = tid; w < wpad; w += MG_NT) Vnew[w] = 0;   // (this path scatters with atomicOr)
				__syncthreads();
				const uint8_t *rec = P.img + roff[r0];
				const uint32_t l = mg_ld_u32_unaligned(rec + 1), zt = m - n1;
				const uint8_t *rle = rec + 5;
				__shared__ uint32_t carry[4];
				if (tid == 0) { carry[0] = 0; carry[1] = 0; }
				for (uint32_t cb = 0; cb < l; cb += MG_RAW) {
					const uint32_t nn = l - cb < (uint32_t)MG_RAW ? l - cb : (uint32_t)MG_RAW;
					for (uint32_t i = tid; i < nn; i += MG_NT) raw[i] = rle[cb + i];
					__syncthreads();
					const uint32_t cs = carry[0];
					if (warp == 0) {
						uint32_t tot = cs, ones = carry[1];
						for (uint32_t base = 0; base < nn; base += 32) {
							const uint32_t i = base + lane;
							const uint32_t c = i < nn ? raw[i] : 0u;
							const uint32_t L = mg_rle_len(c), b = c & 1u, L1 = b ? L : 0u;
							uint32_t x = L, y = L1;
							#pragma unroll
							for (int d = 1; d < 32; d <<= 1) {
								const uint32_t tx = __shfl_up_sync(0xffffffffu, x, d), ty = __shfl_up_sync(0xffffffffu, y, d);
								if (lane >= d) { x += tx; y += ty; }
							}
							const uint32_t start = tot + x - L, ones_before = ones + y - L1;
							if (i < nn) { ts[i] = start; td[i] = b ? (int32_t)(zt - (start - ones_before)) : -(int32_t)ones_before; }
							tot += __shfl_sync(0xffffffffu, x, 31);
							ones += __shfl_sync(0xffffffffu, y, 31);
						}
						if (lane == 0) { carry[2] = tot; carry[3] = ones; }
					}
					__syncthreads();
					const uint32_t ce = carry[2];
					// source words overlapping the rank range [cs, ce) of this piece
					const int w_lo = (int)(cs >> 5), w_hi = (int)((ce + 31) >> 5);
					for (int w = w_lo + tid; w < w_hi && w < words; w += MG_NT) {
						const uint32_t bits = Vold[w];
						if (bits == 0) continue;
						const uint32_t pos = (uint32_t)w * 32;
						uint32_t lo = 0;
						for (uint32_t len = nn; len > 1;) { const uint32_t half = len >> 1; lo += ts[lo + half] <= pos ? half : 0u; len -= half; }
						for (uint32_t j = lo; j < nn; ++j) {
							const uint32_t s = ts[j], e = j + 1 < nn ? ts[j + 1] : ce;
							if (s >= pos + 32) break;
							const uint32_t a = s > pos ? s : pos, b = e < pos + 32 ? e : pos + 32;
							if (b <= a) continue;
							const uint32_t piece = (bits >> (a - pos)) & (b - a == 32 ? 0xffffffffu : ((1u << (b - a)) - 1u));
							if (piece == 0) continue;
							const uint32_t dst = a + (uint32_t)td[j], dw = dst >> 5, db = dst & 31;
							atomicOr(&Vnew[dw], piece << db);
							if (db && (piece >> (32 - db))) atomicOr(&Vnew[dw + 1], piece >> (32 - db));
							if (dst >= zt) cnt += __popc(piece);
						}
					}
					__syncthreads();
					if (tid == 0) { carry[0] = carry[2]; carry[1] = carry[3]; }
					__syncthreads();
				}
			}
			if (n1 != 0 && n1 != m) {
				#pragma unroll
				for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
				if (lane == 0 && cnt) atomicAdd(&r_cnt[r], cnt);
				__syncthreads();                             // all bits have landed in Vnew
				uint32_t *t = Vold; Vold = Vnew; Vnew = t;
			}
		}
		__syncthreads();
		if (tid < nr) {
			const long long arow = blk_row + r0 + tid;
			if (arow >= P.row_lo && arow < P.row_hi) P.n0g[(size_t)(arow - P.row_lo) * P.n_vec + g] = r_cnt[tid];
			r_cnt[tid] = 0;
		}
		__syncthreads();
		r0 += nr;
	}
}

// Segment vectors: the block's start vector pushed through the composite maps of its row groups (compose.cu), stored in
// front of every segment of seg_groups groups, so that the segments of a block can be walked row by row side by side.
// A composite is a list of pieces (start, translation) ordered by start: every thread moves a contiguous stretch of
// source words, piece by piece, with shared-memory atomicOr (the pieces of the next group are fetched meanwhile).
constexpr int MS_NT = 1024;

// One step of a vector through a composite map staged in shared memory (cs/cd: piece starts and translations, np pieces).  Every
// thread owns the words [w_lo, w_hi).  gather: the pieces are cut in the coordinates of Vnew (the thread assembles its own output
// words); otherwise in those of Vold (the thread scatters its source words with atomicOr; Vnew must be zero).
__device__ __forceinline__ void mg_advance(const uint32_t *Vold, uint32_t *Vnew, const uint32_t *cs, const int32_t *cd, uint32_t np, uint32_t m,
                                           int w_lo, int w_hi, bool gather)
{
	if (w_lo >= w_hi) return;
	uint32_t cur = (uint32_t)w_lo * 32;
	const uint32_t end = (uint32_t)w_hi * 32 < m ? (uint32_t)w_hi * 32 : m;
	uint32_t k = 0;
	for (uint32_t len = np; len > 1;) { const uint32_t half = len >> 1; k += cs[k + half] <= cur ? half : 0u; len -= half; }
	uint32_t next = k + 1 < np ? cs[k + 1] : m;
	uint32_t delta = (uint32_t)cd[k];
	// ONE loop over the stretches between consecutive piece starts and word boundaries (a loop over words with a loop over pieces
	// inside makes the lanes of a warp wait for each other at every word)
	if (gather) {
		uint32_t acc = 0;
		while (cur < end) {
			const uint32_t wend = (cur | 31u) + 1u;
			uint32_t stop = next < wend ? next : wend;
			if (stop > end) stop = end;
			const uint32_t take = stop - cur, src = cur + delta, sw = src >> 5, sb = src & 31u;
			if (src < m) {                                          // (a well-formed map never leaves [0, m))
				const uint32_t word = __funnelshift_r(Vold[sw], Vold[sw + 1], sb);   // bits src .. src+31 (the vectors carry spare zero words)
				acc |= (word & (take == 32 ? 0xffffffffu : ((1u << take) - 1u))) << (cur & 31u);
			}
			cur = stop;
			if (cur == wend || cur == end) { Vnew[(cur - 1u) >> 5] = acc; acc = 0; }
			if (cur >= next) { ++k; next = k + 1 < np ? cs[k + 1] : m; delta = (uint32_t)cd[k]; }
		}
	} else {
		while (cur < end) {
			const uint32_t wend = (cur | 31u) + 1u;
			uint32_t stop = next < wend ? next : wend;
			if (stop > end) stop = end;
			const uint32_t take = stop - cur;
			const uint32_t piece = (Vold[cur >> 5] >> (cur & 31u)) & (take == 32 ? 0xffffffffu : ((1u << take) - 1u));
			if (piece) {
				const uint32_t dst = cur + delta, dw = dst >> 5, db = dst & 31u;
				if (dst < m) {                                      // (a well-formed map never leaves [0, m))
					atomicOr(&Vnew[dw], piece << db);
					if (db && (piece >> (32 - db))) atomicOr(&Vnew[dw + 1], piece >> (32 - db));
				}
			}
			cur = stop;
			if (cur >= next) { ++k; next = k + 1 < np ? cs[k + 1] : m; delta = (uint32_t)cd[k]; }
		}
	}
}

__global__ void __launch_bounds__(MS_NT) pbwt_marginal_seed_kernel(const MarginalParams P)
{
	extern __shared__ __align__(16) uint8_t sm[];
	const int words = (P.m + 31) / 32, wpad = (words + 4 + 3) & ~3;
	uint32_t *V0 = (uint32_t*)sm, *V1 = V0 + wpad;
	uint32_t *cs = V1 + wpad;                        // [COMP_CAP] piece starts
	int32_t *cd = (int32_t*)(cs + COMP_CAP);         // [COMP_CAP] translations
	__shared__ int s_bad;
	constexpr int PER = COMP_CAP / MS_NT;
	const int tid = threadIdx.x;
	const int bi = blockIdx.x, blk = P.blk_list ? P.blk_list[bi] : P.blk_first + bi, g = blockIdx.y;
	if (P.blk_ok && !P.blk_ok[blk]) return;
	const uint32_t m = (uint32_t)P.m;
	const long long blk_row = P.blk_row0 + ((long long)blk << P.shift);
	int rows = P.rows_in_blk[blk];
	if (blk_row + rows > P.row_hi) rows = (int)(P.row_hi - blk_row);
	const int seg_rows = P.seg_groups * COMP_K;
	const int n_seg_used = rows > 0 ? (rows + seg_rows - 1) / seg_rows : 0;
	const int g_end = n_seg_used > 1 ? (n_seg_used - 1) * P.seg_groups : 0;   // groups in front of the last segment start
	const size_t slot0 = (size_t)blk * P.n_grp;
	if (tid == 0) s_bad = 0;
	__syncthreads();
	for (int i = tid; i < g_end; i += MS_NT) if (P.comp_n[slot0 + i] <= 0 || P.comp_n[slot0 + i] > COMP_CAP) s_bad = 1;
	__syncthreads();
	if (s_bad || g_end == 0) { if (tid == 0) P.seg_ok[(size_t)bi * P.n_vec + g] = s_bad ? 0 : 1; return; }
	if (tid == 0) P.seg_ok[(size_t)bi * P.n_vec + g] = 1;
	mg_snapshot_vector<MS_NT>(P, blk, g, wpad, V0, V1);
	uint32_t *Vold = V0, *Vnew = V1;
	uint32_t ps[PER]; int32_t pd[PER]; int np_next = 0;
	auto fetch = [&](int gg) {
		np_next = P.comp_n[slot0 + gg];
		const uint32_t *s = P.comp_start + (slot0 + gg) * COMP_CAP; const int32_t *d = P.comp_delta + (slot0 + gg) * COMP_CAP;
		#pragma unroll
		for (int j = 0; j < PER; ++j) { const int i = tid + j * MS_NT; if (i < np_next) { ps[j] = s[i]; pd[j] = d[i]; } }
	};
	fetch(0);
	uint32_t *out = P.vseg + ((size_t)bi * P.n_vec + g) * P.seg_slots * (size_t)wpad;
	const int per = (words + MS_NT - 1) / MS_NT;
	const int w_lo = tid * per, w_hi = w_lo + per < words ? w_lo + per : words;
	for (int gg = 0; gg < g_end; ++gg) {
		const uint32_t np = (uint32_t)np_next;
		#pragma unroll
		for (int j = 0; j < PER; ++j) { const uint32_t i = tid + j * MS_NT; if (i < np) { cs[i] = ps[j]; cd[i] = pd[j]; } }
		for (int w = tid; w < wpad; w += MS_NT) Vnew[w] = 0;
		if (gg + 1 < g_end) fetch(gg + 1);
		__syncthreads();
		// two-sided maps (compose.cu): the groups of the second half carry the INVERSE map, whose pieces are cut in the coordinates
		// behind the group -- the vector is gathered through it (every thread owns its output words) instead of scattered
		const bool inv_map = gg >= comp_first_inverse(blk, P.n_blk_res, P.rows_in_blk[blk], 1 << P.shift, P.n_grp, P.two_sided);
		mg_advance(Vold, Vnew, cs, cd, np, m, w_lo, w_hi, inv_map);
		__syncthreads();
		uint32_t *t = Vold; Vold = Vnew; Vnew = t;
		if ((gg + 1) % P.seg_groups == 0) {
			uint32_t *dst = out + (size_t)((gg + 1) / P.seg_groups) * wpad;
			for (int w = tid; w < wpad; w += MS_NT) dst[w] = w < words ? Vold[w] : 0u;
		}
	}
}

// The vectors margpiece.cu works from: one in front of every 32-row group the scan reaches and one behind every full one of them,
// as look-up records {64 bits of the vector, members of the group in front of them}.  Two CTAs per block where the block is full,
// scanned to its end and followed by a resident block (the condition under which compose.cu builds inverse maps for the second half
// of the groups): one walks forward from the block's snapshot to the middle, the other BACKWARD from the next block's snapshot --
// an inverse map, cut in the coordinates behind its group, scatters a vector from behind the group to in front of it.
// seg_ok must have been set to 1 on the stream; a missing composite map clears it.
__global__ void __launch_bounds__(MS_NT, 2) pbwt_marginal_dense_seed_kernel(const MarginalParams P)
{
	extern __shared__ __align__(16) uint8_t sm[];
	const int words = (P.m + 31) / 32, wpad = (words + 4 + 3) & ~3;
	uint32_t *V0 = (uint32_t*)sm, *V1 = V0 + wpad;
	uint32_t *cs = V1 + wpad;                        // [COMP_CAP] piece starts
	int32_t *cd = (int32_t*)(cs + COMP_CAP);         // [COMP_CAP] translations
	__shared__ int s_bad;
	__shared__ uint32_t s_warp[MS_NT / 32];
	constexpr int PER = COMP_CAP / MS_NT;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int bi = (int)(blockIdx.x >> 1), side = (int)(blockIdx.x & 1), blk = P.blk_first + bi, g = blockIdx.y;
	if (P.blk_ok && !P.blk_ok[blk]) return;
	const uint32_t m = (uint32_t)P.m;
	const int BS = 1 << P.shift, n_grp = P.n_grp;
	const long long blk_row = P.blk_row0 + ((long long)blk << P.shift);
	const int rows_all = P.rows_in_blk[blk];
	int rows = rows_all;
	if (blk_row + rows > P.row_hi) rows = (int)(P.row_hi - blk_row);
	if (rows <= 0) return;
	const int n_full = rows_all / COMP_K, n_used = (rows + COMP_K - 1) / COMP_K;
	const int n_adv = n_full < n_used ? n_full : n_used;             // vectors 0 .. n_adv are wanted
	const int first_inv = comp_first_inverse(blk, P.n_blk_res, rows_all, BS, n_grp, P.two_sided);
	const bool two = first_inv < n_grp && rows == BS && n_grp >= 4;
	if (side == 1 && !two) return;
	// the groups this CTA crosses, in order: forward g_a, g_a + 1, ... g_b - 1, or backward g_a - 1, g_a - 2, ... g_b
	const int g_a = side ? n_grp : 0, g_b = side ? n_grp / 2 + 1 : (two ? n_grp / 2 : n_adv);
	const int n_cross = side ? g_a - g_b : g_b - g_a;
	const size_t slot0 = (size_t)blk * n_grp;
	if (tid == 0) s_bad = 0;
	__syncthreads();
	for (int i = tid; i < n_cross; i += MS_NT) { const int gg = side ? g_a - 1 - i : g_a + i; if (P.comp_n[slot0 + gg] <= 0 || P.comp_n[slot0 + gg] > COMP_CAP) s_bad = 1; }
	__syncthreads();
	if (s_bad) { if (tid == 0) P.seg_ok[(size_t)bi * P.n_vec + g] = 0; return; }
	mg_snapshot_vector<MS_NT>(P, side ? blk + 1 : blk, g, wpad, V0, V1);
	uint32_t *Vold = V0, *Vnew = V1;
	uint32_t ps[PER]; int32_t pd[PER]; int np_next = 0;
	auto fetch = [&](int gg) {
		np_next = P.comp_n[slot0 + gg];
		const uint32_t *s = P.comp_start + (slot0 + gg) * COMP_CAP; const int32_t *d = P.comp_delta + (slot0 + gg) * COMP_CAP;
		#pragma unroll
		for (int j = 0; j < PER; ++j) { const int i = tid + j * MS_NT; if (i < np_next) { ps[j] = s[i]; pd[j] = d[i]; } }
	};
	if (n_cross > 0) fetch(side ? g_a - 1 : g_a);
	const int W64 = (int)marginal_rec_words64_dev(P.m);
	uint4 *out = P.vrec + ((size_t)bi * P.n_vec + g) * P.seg_slots * (size_t)W64;
	const int per64 = (W64 + MS_NT - 1) / MS_NT;
	const int r_lo = tid * per64, r_hi = r_lo + per64 < W64 ? r_lo + per64 : W64;
	// the vector as look-up records (all threads; V complete and not written meanwhile)
	auto store = [&](const uint32_t *V, int slot) {
		uint32_t s = 0;
		for (int w = r_lo; w < r_hi; ++w) s += (uint32_t)(__popc(V[2 * w]) + __popc(V[2 * w + 1]));
		uint32_t x = s;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
		if (lane == 31) s_warp[warp] = x;
		__syncthreads();
		uint32_t run = x - s;
		{
			const uint32_t v = lane < warp ? s_warp[lane] : 0u;       // (32 warps: one lane per warp total)
			uint32_t y = v;
			#pragma unroll
			for (int d = 16; d; d >>= 1) y += __shfl_xor_sync(0xffffffffu, y, d);
			run += y;
		}
		uint4 *dst = out + (size_t)slot * W64;
		for (int w = r_lo; w < r_hi; ++w) {
			const uint32_t a = V[2 * w], b = V[2 * w + 1];
			dst[w] = make_uint4(a, b, run, 0u);
			run += (uint32_t)(__popc(a) + __popc(b));
		}
	};
	__syncthreads();
	store(Vold, g_a);
	const int per = (words + MS_NT - 1) / MS_NT;
	const int w_lo = tid * per, w_hi = w_lo + per < words ? w_lo + per : words;
	for (int i = 0; i < n_cross; ++i) {
		const int gg = side ? g_a - 1 - i : g_a + i;
		const uint32_t np = (uint32_t)np_next;
		#pragma unroll
		for (int j = 0; j < PER; ++j) { const uint32_t q = tid + j * MS_NT; if (q < np) { cs[q] = ps[j]; cd[q] = pd[j]; } }
		for (int w = tid; w < wpad; w += MS_NT) Vnew[w] = 0;
		if (i + 1 < n_cross) fetch(side ? gg - 1 : gg + 1);
		__syncthreads();                                              // (also: everybody is done with s_warp and with reading Vnew's old content)
		// forward through a forward map and backward through an inverse map: the pieces are cut where the vector comes from -> scatter;
		// forward through an inverse map (a block that is not walked from both ends): gather
		mg_advance(Vold, Vnew, cs, cd, np, m, w_lo, w_hi, !side && gg >= first_inv);
		__syncthreads();
		uint32_t *t = Vold; Vold = Vnew; Vnew = t;
		store(Vold, side ? gg : gg + 1);
	}
}

static size_t marginal_smem_raw(int m, int raw)
{
	const int words = (m + 31) / 32, wpad = (words + 4 + 3) & ~3;
	return (size_t)wpad * 8 + (size_t)raw * 8 + raw + 16 + MG_TMAX * 24 + (size_t)raw * 2 + 64;
}

size_t marginal_seg_words(int m) { return (size_t)(((m + 31) / 32 + 4 + 3) & ~3); }

size_t marginal_smem_bytes(int m)
{
	return marginal_smem_raw(m, MG_RAW);
}

cudaError_t launch_marginal_seed(const MarginalParams &P, int n_blk, cudaStream_t st)
{
	const int words = (P.m + 31) / 32, wpad = (words + 4 + 3) & ~3;
	const size_t smem_a = (size_t)wpad * 8 + (size_t)COMP_CAP * 8;
	cudaError_t e = cudaFuncSetAttribute(pbwt_marginal_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
	if (e != cudaSuccess) return e;
	pbwt_marginal_seed_kernel<<<dim3(n_blk, P.n_vec, 1), MS_NT, smem_a, st>>>(P);
	return cudaGetLastError();
}

cudaError_t launch_marginal_dense_seed(const MarginalParams &P, int n_blk, cudaStream_t st)
{
	const int words = (P.m + 31) / 32, wpad = (words + 4 + 3) & ~3;
	const size_t smem_a = (size_t)wpad * 8 + (size_t)COMP_CAP * 8;
	cudaError_t e = cudaFuncSetAttribute(pbwt_marginal_dense_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
	if (e != cudaSuccess) return e;
	pbwt_marginal_dense_seed_kernel<<<dim3(2 * n_blk, P.n_vec, 1), MS_NT, smem_a, st>>>(P);
	return cudaGetLastError();
}

// the row loop alone (the vectors in front of the segments are there, or n_seg == 1: every block from its snapshot)
cudaError_t launch_marginal_rows(const MarginalParams &P, int n_blk, cudaStream_t st)
{
	cudaError_t e;
	if (P.n_seg > 1) {
		const size_t smem_b = marginal_smem_raw(P.m, MG_RAW_SEG);
		e = cudaFuncSetAttribute(pbwt_marginal_kernel<MG_NT_SEG, MG_RAW_SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
		if (e != cudaSuccess) return e;
		pbwt_marginal_kernel<MG_NT_SEG, MG_RAW_SEG><<<dim3(n_blk * P.n_seg, P.n_vec, 1), MG_NT_SEG, smem_b, st>>>(P);
		return cudaGetLastError();
	}
	const size_t smem = marginal_smem_bytes(P.m);
	e = cudaFuncSetAttribute(pbwt_marginal_kernel<MG_NT, MG_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	dim3 grid(n_blk, P.n_vec, 1);
	pbwt_marginal_kernel<MG_NT, MG_RAW><<<grid, MG_NT, smem, st>>>(P);
	return cudaGetLastError();
}

cudaError_t launch_marginal(const MarginalParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0 || P.n_vec <= 0) return cudaSuccess;
	if (P.n_seg > 1) { // segment vectors first, then every segment on its own (smaller CTAs, three to an SM)
		const cudaError_t e = launch_marginal_seed(P, n_blk, st);
		if (e != cudaSuccess) return e;
	}
	return launch_marginal_rows(P, n_blk, st);
}

} // namespace b200
