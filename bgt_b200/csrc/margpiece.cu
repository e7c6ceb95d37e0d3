// margpiece.cu -- per-group plane-0 marginals without touching the rank vector row by row (r2; replaces the row loop of
// marginal.cu for everything it can hold, marginal.cu stays as the general fall-back).
//
// marginal.cu pushes one bit per rank ("the column at this rank belongs to the group") through every row's stable partition
// (pbwt.c:79-88): m/32 words per row and group.  But the partition of a row with R runs only cuts the rank axis in R places:
// the composition of the partitions of k consecutive rows is a piecewise translation with < k*R pieces.  So keep the MAP
// instead of the vector:
//
//     list_k = { (d_i, delta_i) } sorted by d:  the column at rank x of row k, d_i <= x < d_{i+1}, sat at rank x + delta_i
//                                                 of the reference row (where a bit vector with prefix counts is at hand)
//
// and the group's number of ones of row k is a prefix query on the row behind it: the 1-runs of row k land, in order, on
// ranks [zt_k, m) of row k+1 (zt_k = m - n1_k), hence
//
//     ones_k(group) = sum over the pieces of list_{k+1} inside [zt_k, m) of  P(d_{i+1} + delta_i) - P(max(d_i, zt_k) + delta_i)
//
// with P(x) = members of the group among the first x ranks of the reference row -- a few dozen look-ups per row instead of
// m/32 words (or, if the head [0, zt_k) holds fewer pieces, the group's size minus the same sum over the head).  Composing the
// list with one more row is a segmented copy: every run takes the pieces it overlaps, clipped, to its landing place.
//
// Reference rows come from the seed kernel of marginal.cu (the block's start vector pushed through the composite maps of
// compose.cu), stored in front of EVERY 32-row group as look-up records {64 bits, members in front of them}
// (pbwt_marginal_dense_seed_kernel).  A group is worked on from both ends, by two CTAs ("chains"): rows 0..15 forward from the vector in front of
// the group (list_k: row k -> group start), rows 31..16 backward from the vector behind it (list_k: row k -> group end), so that
// no list holds the cuts of more than 16 rows.  The look-ups of a row are issued when its list is ready and consumed one row
// later: their latency hides behind the next row's composition.
//
// A chain that does not fit the small configuration (more pieces or runs than its tables hold) is queued for the long
// configuration, which also walks the ragged last group of a file (no vector behind it: all its rows forward).  What does not
// fit there either -- and blocks without composite maps -- raises the block's flag in blk_fail and is redone by marginal.cu's row
// loop (queued behind these kernels; its CTAs exit at once for blocks without the flag).
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

constexpr int MP_NT = 32;                                     // a chain is one warp's work
constexpr int MP_HALF = COMP_K / 2;
constexpr int MP_RETRY_CTAS = 296;

struct MpSmall { static constexpr int ROWS = MP_HALF, PC = 1408, RMAX = 256; };    // 27.8 KB of shared memory: eight chains to an SM
struct MpLong  { static constexpr int ROWS = COMP_K,  PC = 3584, RMAX = 2048; };

__device__ __forceinline__ uint32_t mp_rle_len(uint32_t c) { const uint32_t v = c >> 1; return (v & 15u) << ((v >> 4) << 2); }

__device__ __forceinline__ uint32_t mp_ld_u32_unaligned(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	const uint32_t lo = w[0];
	if (sh == 0) return lo;
	return __funnelshift_r(lo, w[1], sh);
}

// last index i in [0, n) with a[i] <= x (a[0] <= x)
template<typename T>
__device__ __forceinline__ int mp_last_le(const T *a, int n, uint32_t x)
{
	int k = 0;
	for (int len = n; len > 1;) { const int half = len >> 1; k += (uint32_t)a[k + half] <= x ? half : 0; len -= half; }
	return k;
}

__device__ __forceinline__ uint32_t mp_rank(const uint4 r, uint32_t x)   // members among the first x ranks, from the record of word x >> 6
{
	const unsigned long long bits = (unsigned long long)r.x | (unsigned long long)r.y << 32;
	return r.z + (uint32_t)__popcll(bits & ((1ull << (x & 63u)) - 1ull));
}

// ------------------------------------------------------------------------------------------------ one chain

struct MpChain { int bi, g, gg, back, n_steps, row_first; };

// Runs one chain: ONE WARP, no CTA-wide synchronisation (every lane of the warp calls it with the same arguments; sm is the warp's
// own shared memory).  False: the chain does not fit this configuration.
template<typename CFG>
__device__ bool mp_chain(const MarginalParams &P, const MpChain c, uint8_t *sm)
{
	constexpr int ROWS = CFG::ROWS, PC = CFG::PC, RMAX = CFG::RMAX;
	uint32_t *d = (uint32_t*)sm;                                  // piece lists, two buffers
	int32_t *l = (int32_t*)(d + PC + 4);
	uint32_t *nd = (uint32_t*)(l + PC + 4);
	int32_t *nl = (int32_t*)(nd + PC + 4);
	uint32_t *U = (uint32_t*)(nl + PC + 4);                       // run table of the row at hand: start of every run on the list's axis (+ sentinel m) ...
	uint32_t *V = U + RMAX + 4;                                   // ... and where it goes on the other side of the row
	uint32_t *hu = V + RMAX + 4;                                  // per run in OUTPUT order: its start on the list's axis, ...
	int32_t *hw = (int32_t*)(hu + RMAX + 4);                      // ... the shift to the other side of the row, ...
	uint16_t *hb = (uint16_t*)(hw + RMAX + 4);                    // ... the first piece it overlaps, ...
	uint16_t *off = hb + RMAX + 4;                                // ... the pieces in front of it in the new list
	int *r_c = (int*)(off + RMAX + 4);                            // [ROWS] sums

	const int lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t m = (uint32_t)P.m;
	const int BS = 1 << P.shift;
	const int blk = P.blk_first + c.bi, g = c.g;
	const bool back = c.back != 0;
	const int n_steps = c.n_steps;
	const int W64 = (int)((m + 63) / 64) + 1;
	const uint4 *rec = P.vrec + ((((size_t)c.bi * P.n_vec + g) * P.seg_slots) + (size_t)(c.gg + (back ? 1 : 0))) * (size_t)W64;
	const uint64_t *roff = P.rowoff + (size_t)blk * (BS + 1);

	// what lane k knows about the row of step k
	uint32_t md_nr = 0, md_n1 = 0, md_len = 0;
	const uint8_t *md_rle = P.img;
	if (lane < n_steps) {
		const int r = back ? c.row_first - lane : c.row_first + lane;
		md_nr = P.nrun[(size_t)blk * BS + r]; md_n1 = P.n1[((size_t)blk * BS + r) * 2];
		const uint8_t *rcd = P.img + roff[r];
		md_len = mp_ld_u32_unaligned(rcd + 1);
		md_rle = rcd + 5;
	}
	__syncwarp();                                                  // (the previous chain of this warp is done with the shared memory)
	if (lane < ROWS) r_c[lane] = 0;
	if (lane == 0) { d[0] = 0; l[0] = 0; d[1] = m; }
	int np = 1;
	uint32_t tmask = 0;                                            // rows whose sum is "group size minus head" (or all ones)
	uint4 qT = make_uint4(0, 0, 0, 0);
	if (lane == 0) qT = rec[m >> 6];                               // the group's size, consumed at the end

	// the first 128 bytes of the next row's RLE, four per lane, fetched one row ahead
	uint32_t pf_lo = 0, pf_hi = 0;
	auto prefetch = [&](int k) {
		const uint32_t len = __shfl_sync(0xffffffffu, md_len, k);
		const uint8_t *rle = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)md_rle, k);
		pf_lo = pf_hi = 0;
		if (4u * lane < len) {
			const uintptr_t a = (uintptr_t)(rle + 4 * lane);
			const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
			pf_lo = w[0];
			if (a & 3) pf_hi = w[1];
		}
	};
	prefetch(0);

	// look-ups in flight: one pair per lane, consumed one row later
	uint4 qa = make_uint4(0, 0, 0, 0), qe = qa;
	uint32_t xa = 0, xe = 0;
	int qk = -1, qsign = 0;
	auto consume = [&]() {
		if (qk >= 0) {
			const int v = (int)(mp_rank(qe, xe) - mp_rank(qa, xa));
			if (v) atomicAdd(&r_c[qk], qsign * v);
			qk = -1;
		}
	};
	// ones of the group in the row (step k) whose zeros number zt, from the list of the row BEHIND it; i0 = the piece that holds rank zt
	auto tail_query = [&](const uint32_t *dd, const int32_t *ll, int n, int i0, uint32_t zt, int k) {
		const bool tail = n - i0 <= i0 + 1;
		const int ia = tail ? i0 : 0, ib = tail ? n - 1 : i0;
		const uint32_t xlo = tail ? zt : 0u, xhi = tail ? m : zt;
		if (!tail) tmask |= 1u << k;
		for (int i = ia + lane; i <= ib; i += 32) {
			const uint32_t a = dd[i] > xlo ? dd[i] : xlo, e = dd[i + 1] < xhi ? dd[i + 1] : xhi;
			if (e > a) {
				consume();
				xa = a + (uint32_t)ll[i]; xe = e + (uint32_t)ll[i];
				qa = rec[xa >> 6]; qe = rec[xe >> 6];
				qk = k; qsign = tail ? 1 : -1;
			}
		}
	};

	for (int k = 0; k < n_steps; ++k) {
		const uint32_t nr = __shfl_sync(0xffffffffu, md_nr, k), n1 = __shfl_sync(0xffffffffu, md_n1, k);
		const uint32_t len = __shfl_sync(0xffffffffu, md_len, k);
		const uint8_t *rle = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)md_rle, k);
		uint32_t R = nr & 0x7fffffffu;
		const uint32_t b0 = nr >> 31, zt = m - n1;
		if (n1 == 0 || n1 == m) R = 1;                              // nothing moves (a corrupt row counts as all-REF, like everywhere)
		const uint32_t nzr = b0 ? R >> 1 : (R + 1) >> 1;
		const bool more = k + 1 < n_steps;
		const uint32_t w_lo = pf_lo, w_hi = pf_hi;
		if (more) prefetch(k + 1);
		if (R <= 1) { if (zt == 0) tmask |= 1u << k; continue; }   // all zeros: no ones; all ones: every member; nothing moves
		if (R > (uint32_t)RMAX) return false;

		// ---- run table: where every run starts and where it lands (pbwt.c:79-88: 0-runs in front in order, 1-runs behind the zt zeros
		// in order); forward lists are cut on the row's own axis and carried to the landing axis, backward lists the other way round.
		// Four RLE bytes per lane and pass.
		{
			uint32_t tot = 0, ones = 0, jn = 0, prev_bit = 2;
			for (uint32_t base = 0; base < len; base += 128) {
				const uint32_t i0 = base + 4u * lane;
				uint32_t word = 0;
				if (base == 0) word = __funnelshift_r(w_lo, w_hi, (uint32_t)((uintptr_t)(rle + 4 * lane) & 3) * 8);
				else if (i0 < len) word = mp_ld_u32_unaligned(rle + i0);
				uint32_t L[4], B[4], sx = 0, sy = 0;
				#pragma unroll
				for (int q = 0; q < 4; ++q) {
					const uint32_t cc = i0 + q < len ? (word >> (8 * q)) & 0xffu : 0u;
					L[q] = mp_rle_len(cc); B[q] = cc & 1u;
					sx += L[q]; sy += B[q] ? L[q] : 0u;
				}
				uint32_t x = sx, y = sy;
				#pragma unroll
				for (int s = 1; s < 32; s <<= 1) {
					const uint32_t tx = __shfl_up_sync(0xffffffffu, x, s), ty = __shfl_up_sync(0xffffffffu, y, s);
					if (lane >= s) { x += tx; y += ty; }
				}
				// the bit of the last non-empty byte in front of this lane's bytes
				uint32_t lastb = 0; bool hasv = false;
				#pragma unroll
				for (int q = 0; q < 4; ++q) if (L[q]) { hasv = true; lastb = B[q]; }
				const uint32_t vm = __ballot_sync(0xffffffffu, hasv), bm = __ballot_sync(0xffffffffu, lastb != 0);
				const uint32_t below = vm & lt;
				uint32_t pb = below ? (bm >> (31 - __clz(below))) & 1u : prev_bit;
				bool st[4];
				#pragma unroll
				for (int q = 0; q < 4; ++q) { st[q] = L[q] > 0 && pb != B[q]; if (L[q]) pb = B[q]; }
				uint32_t j = jn, n_st = 0;
				#pragma unroll
				for (int q = 0; q < 4; ++q) { const uint32_t bq = __ballot_sync(0xffffffffu, st[q]); j += __popc(bq & lt); n_st += __popc(bq); }
				uint32_t pos = tot + x - sx, ob = ones + y - sy;
				#pragma unroll
				for (int q = 0; q < 4; ++q) {
					if (st[q]) {
						const uint32_t land = B[q] ? zt + ob : pos - ob;
						const uint32_t zi = (B[q] ? nzr : 0u) + (j >> 1);     // index in landing order (runs alternate)
						if (j < R) {
							if (back) { U[zi] = land; V[zi] = pos; }
							else { U[j] = pos; V[j] = land; }
						}
						++j;
					}
					pos += L[q]; ob += B[q] ? L[q] : 0u;
				}
				jn += n_st;
				if (vm) prev_bit = (bm >> (31 - __clz(vm))) & 1u;
				tot += __shfl_sync(0xffffffffu, x, 31);
				ones += __shfl_sync(0xffffffffu, y, 31);
			}
			if (jn != R || tot != m) return false;                     // (uniform)
			if (lane == 0) U[R] = m;
		}
		__syncwarp();
		// ---- (1) per run (in the order of the list's axis): the pieces it overlaps.  A pass takes 31 runs: lane j learns where its run
		// ends from lane j+1
		int i0b = 0;
		for (uint32_t t0 = 0; t0 < R; t0 += 31u) {
			const uint32_t t = t0 + lane;
			int a = np; bool exact = true;
			if (t < R) { a = mp_last_le(d, np, U[t]); exact = d[a] == U[t]; }
			const int a_next = __shfl_down_sync(0xffffffffu, a, 1);
			const bool exact_next = __shfl_down_sync(0xffffffffu, (int)exact, 1) != 0;
			if (lane < 31 && t < R) {
				const int b = exact_next ? a_next - 1 : a_next;          // (t + 1 == R: a_next = np, "exact")
				uint32_t s;                                              // index of the run in output order
				if (!back) { const uint32_t bit = b0 ^ (t & 1u); s = (bit ? nzr : 0u) + (t >> 1); }
				else s = t < nzr ? 2u * t + b0 : 2u * (t - nzr) + (b0 ^ 1u);
				off[s] = (uint16_t)(b - a + 1);
				hb[s] = (uint16_t)a; hu[s] = U[t]; hw[s] = (int32_t)(U[t] - V[t]);   // the run as the copy wants it: first piece, start, shift
				if (t == nzr) i0b = a;                                   // (backward: U[nzr] == zt)
			}
		}
		i0b = __shfl_sync(0xffffffffu, i0b, (int)(nzr % 31u));
		__syncwarp();
		// ---- (2) exclusive prefix over the runs in output order
		int np2;
		{
			const uint32_t per = (R + 31) / 32, t0 = lane * per, t1 = t0 + per < R ? t0 + per : R;
			uint32_t s = 0;
			for (uint32_t t = t0; t < t1; ++t) s += off[t];
			uint32_t x = s;
			#pragma unroll
			for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, x, dd); if (lane >= dd) x += v; }
			const uint32_t total = __shfl_sync(0xffffffffu, x, 31);
			if (total > (uint32_t)PC) return false;                    // (uniform)
			uint32_t run = x - s;
			for (uint32_t t = t0; t < t1; ++t) { const uint32_t cc = off[t]; off[t] = (uint16_t)run; run += cc; }
			if (lane == 0) off[R] = (uint16_t)total;
			np2 = (int)total;
		}
		__syncwarp();
		if (back) {
			tail_query(d, l, np, i0b, zt, k);                        // (the list still belongs to row k+1)
			if (!more) break;                                         // (nothing is asked of the list in front of the chain's last row)
		}
		// ---- (3) segmented copy, a contiguous stretch of output slots per lane (odd length: no bank conflicts)
		const int i0f = (int)off[nzr];
		{
			const int C = ((np2 + 31) / 32) | 1;
			int o = lane * C;
			if (o < np2) {
				uint32_t s = (uint32_t)mp_last_le(off, (int)R, (uint32_t)o);
				int nxt = (int)off[s + 1], bi = (int)hb[s] - (int)off[s];
				uint32_t u = hu[s]; int32_t w = hw[s];
				const int o_end = o + C < np2 ? o + C : np2;
				for (; o < o_end; ++o) {
					if (o == nxt) { ++s; u = hu[s]; w = hw[s]; bi = (int)hb[s] - nxt; nxt = (int)off[s + 1]; }
					const uint32_t di = d[bi + o];
					nd[o] = (di > u ? di : u) - (uint32_t)w;
					nl[o] = l[bi + o] + w;
				}
			}
			if (lane == 0) nd[np2] = m;
		}
		__syncwarp();
		{ uint32_t *t1 = d; d = nd; nd = t1; int32_t *t2 = l; l = nl; nl = t2; }
		np = np2;
		if (!back) tail_query(d, l, np, i0f, zt, k);                // (the new list belongs to row k+1; its piece i0f starts at zt)
	}
	consume();
	__syncwarp();
	const uint32_t T = __shfl_sync(0xffffffffu, mp_rank(qT, m), 0);
	if (lane < n_steps) {
		const int r = back ? c.row_first - lane : c.row_first + lane;
		const long long arow = P.blk_row0 + ((long long)blk << P.shift) + r;
		if (arow >= P.row_lo && arow < P.row_hi) P.n0g[(size_t)(arow - P.row_lo) * P.n_vec + g] = r_c[lane] + ((tmask >> lane) & 1u ? (int)T : 0);
	}
	return true;
}

template<typename CFG>
static size_t mp_smem_bytes()
{
	return 4 * (size_t)(CFG::PC + 4) * 4 + 4 * (size_t)(CFG::RMAX + 4) * 4 + 2 * (size_t)(CFG::RMAX + 4) * 2 + (size_t)CFG::ROWS * 4 + 16;
}

// does the chain (block bi, group gg, direction) exist in this scan?  Fills in its rows.
__device__ __forceinline__ bool mp_chain_rows(const MarginalParams &P, MpChain &c, bool ragged)
{
	const int blk = P.blk_first + c.bi;
	const long long blk_row = P.blk_row0 + ((long long)blk << P.shift);
	const int rows_all = P.rows_in_blk[blk];
	int rows = rows_all;
	if (blk_row + rows > P.row_hi) rows = (int)(P.row_hi - blk_row);
	if (rows <= 0 || blk_row + rows_all <= P.row_lo) return false;
	const int n_full = rows_all / COMP_K;
	if (ragged) { c.gg = n_full; c.back = 0; if (c.gg * COMP_K >= rows) return false; }      // the ragged group, if the scan reaches it
	else if (c.gg >= n_full || c.gg * COMP_K >= rows) return false;
	c.row_first = c.gg * COMP_K + (c.back ? COMP_K - 1 : 0);
	c.n_steps = ragged ? rows - c.gg * COMP_K : MP_HALF;
	if (!ragged && !c.back && c.gg * COMP_K + c.n_steps > rows) c.n_steps = rows - c.gg * COMP_K;   // (forward: nothing behind the scan's last row matters)
	// no output row in this chain's reach?
	const long long a = blk_row + (c.back ? c.gg * COMP_K + MP_HALF : c.gg * COMP_K), b = blk_row + (c.back ? (c.gg + 1) * COMP_K : c.gg * COMP_K + c.n_steps);
	return b > P.row_lo && a < P.row_hi;
}

// grid (blocks * groups * 2, vectors): one chain per CTA
__global__ void __launch_bounds__(MP_NT) pbwt_marginal_piece_kernel(const MarginalParams P)
{
	extern __shared__ __align__(16) uint8_t sm[];
	MpChain c;
	const unsigned per_blk = 2u * (unsigned)P.n_grp;
	c.bi = (int)(blockIdx.x / per_blk); c.g = blockIdx.y;
	{ const int r = (int)(blockIdx.x % per_blk); c.gg = r >> 1; c.back = r & 1; }
	const int blk = P.blk_first + c.bi;
	if (P.blk_ok && !P.blk_ok[blk]) return;
	if (!mp_chain_rows(P, c, false)) return;
	uint8_t *fail = P.blk_fail + (size_t)c.bi * P.n_vec + c.g;
	if (!P.seg_ok[(size_t)c.bi * P.n_vec + c.g]) { if (threadIdx.x == 0) *fail = 1; return; }   // no vectors for this block
	if (!mp_chain<MpSmall>(P, c, sm) && threadIdx.x == 0) {      // again with the long configuration
		const int at = atomicAdd(P.retry_n, 1);
		if (at < P.retry_cap) P.retry[at] = make_uint2((uint32_t)c.bi, (uint32_t)c.gg | (uint32_t)c.back << 16 | (uint32_t)c.g << 20);
		else *fail = 1;
	}
}

// grid (blocks + MP_RETRY_CTAS, vectors): the ragged last group of every block, then the chains the small configuration gave up on
__global__ void __launch_bounds__(MP_NT) pbwt_marginal_piece_long_kernel(const MarginalParams P, int n_blk)
{
	extern __shared__ __align__(16) uint8_t sm[];
	if ((int)blockIdx.x < n_blk) {
		MpChain c;
		c.bi = blockIdx.x; c.g = blockIdx.y; c.gg = 0; c.back = 0;
		const int blk = P.blk_first + c.bi;
		if (P.blk_ok && !P.blk_ok[blk]) return;
		if (!mp_chain_rows(P, c, true)) return;
		uint8_t *fail = P.blk_fail + (size_t)c.bi * P.n_vec + c.g;
		if (!P.seg_ok[(size_t)c.bi * P.n_vec + c.g]) { if (threadIdx.x == 0) *fail = 1; return; }
		if (!mp_chain<MpLong>(P, c, sm) && threadIdx.x == 0) *fail = 1;
		return;
	}
	if (blockIdx.y != 0) return;                                  // (the queue holds the chains of all vectors)
	int n = *P.retry_n;
	if (n > P.retry_cap) n = P.retry_cap;
	for (int i = (int)blockIdx.x - n_blk; i < n; i += MP_RETRY_CTAS) {
		const uint2 e = P.retry[i];
		MpChain c;
		c.bi = (int)e.x; c.gg = (int)(e.y & 0xffffu); c.back = (int)((e.y >> 16) & 1u); c.g = (int)(e.y >> 20);
		if (!mp_chain_rows(P, c, false)) continue;
		if (!mp_chain<MpLong>(P, c, sm) && threadIdx.x == 0) P.blk_fail[(size_t)c.bi * P.n_vec + c.g] = 1;
	}
}

size_t marginal_rec_words64(int m) { return (size_t)((m + 63) / 64) + 1; }

// Queues the two chain launches (the vectors are there: launch_marginal_dense_seed).  blk_fail and retry_n must have been cleared on the stream.
cudaError_t launch_marginal_pieces(const MarginalParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0 || P.n_vec <= 0) return cudaSuccess;
	cudaError_t e;
	const size_t sa = mp_smem_bytes<MpSmall>(), sb = mp_smem_bytes<MpLong>();
	if ((e = cudaFuncSetAttribute(pbwt_marginal_piece_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa)) != cudaSuccess) return e;
	if ((e = cudaFuncSetAttribute(pbwt_marginal_piece_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb)) != cudaSuccess) return e;
	pbwt_marginal_piece_kernel<<<dim3((unsigned)((long long)n_blk * P.n_grp * 2), P.n_vec, 1), MP_NT, sa, st>>>(P);
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	pbwt_marginal_piece_long_kernel<<<dim3(n_blk + MP_RETRY_CTAS, P.n_vec, 1), MP_NT, sb, st>>>(P, n_blk);
	return cudaGetLastError();
}

} // namespace b200
