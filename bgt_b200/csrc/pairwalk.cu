// pairwalk.cu -- second phase of the split scan (count-only full-cohort queries, `view -f .. -G`): the plane-0 bit of every
// (column, row) pair that carries a plane-1 bit (missing / other-ALT code; found by plane1.cu).
//
// Reference semantics: bgtm_cal_info (bgt.c:735-757) histograms the 2-bit code a1<<1|a0 of every haplotype; for a pair the
// plane-1 bit is 1, so its code is 3 (other-ALT) if the haplotype's plane-0 bit at that row is 1 and 2 (missing) otherwise.
// The plane-0 bit of column c at row v is the bit of the run that holds c's rank in the PBWT order in front of row v
// (pbwt.c:142-153).  That rank is reached from the nearer of the two snapshots around the row:
//   forward  (target row in the first half of its checkpoint block, or no snapshot behind the block is resident):
//            rank under the block's own 'S' snapshot (pbwt.c:292-301, inverted: rank0) -> one look-up per whole 32-row group in
//            front of the target's group in the group's composite map (compose.cu) -> row by row inside the target's group ->
//            the run of the target row gives the bit;
//   backward (second half): rank under the NEXT block's snapshot = rank behind the block's last row -> one look-up per whole
//            group behind the target's group in the group's INVERSE composite -> rows undone one by one down to the row behind
//            the target -> the rank behind the target row itself tells the bit (ones sit behind the zeros, pbwt.c:79-88).
// Either way a pair crosses at most half a block (<= 128 groups) by composite and <= 31 rows one by one.
//
// Kernel shape: grid = (slices of a block's pair list, blocks), 512 threads, C pairs per thread; pairs are sorted by target
// row, so the pairs of a CTA need (nearly) the same groups.
//   phase A  composite maps stream through shared memory (TMA bulk copies into three stages, mbarrier completion), every
//            warp looks its live pairs up through the map's bucket directory; one CTA barrier per group.
//   phase B  warp-autonomous, no CTA barrier: the 32 pairs of a warp slot are neighbours in the sorted list (a few rows
//            apart), the warp parses the rows of their group itself (RLE bytes straight from L2, two warp scans -> run table
//            in the warp's own shared-memory slice) and the lanes search it; hits go to the per-site counters with one global
//            atomic per (row, group, code) and warp.
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"
#include "pairwalk.cuh"

namespace b200 {

struct PairSmem { uint64_t bar[PW_STAGES]; int f_max, b_min, f_avail, b_avail; };

template<int C>
__global__ void __launch_bounds__(PW_NT, 2) pbwt_pair_kernel(const PairParams P)
{
	extern __shared__ __align__(128) uint8_t pw_sm[];
	__shared__ PairSmem S;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int BS = 1 << P.shift;
	const int blk = P.blk_list ? P.blk_list[blockIdx.y] : P.blk_first + (int)blockIdx.y;
	if (P.blk_ok && !P.blk_ok[blk]) return;
	const int n_pairs = P.qcount[blk];
	const int slice_base = ((int)blockIdx.x + P.slice0) * (PW_NT * C);
	if (slice_base >= n_pairs) return;
	const uint32_t m = (uint32_t)P.m;
	const long long list_at = P.ext_off ? P.ext_off[blk] - P.ext_shift : (long long)blk * P.q_stride;
	const int32_t *qcol = P.qcol + list_at;
	const uint16_t *qrow = P.qrow + list_at;
	const long long blk_row = P.blk_row0 + ((long long)blk << P.shift);
	const int n_grp = (BS + COMP_K - 1) / COMP_K;
	const uint64_t *roff = P.rowoff + (size_t)blk * (BS + 1);
	const uint32_t *n1p = P.n1 + (size_t)blk * BS * 2;
	// first group with an INVERSE map: pairs whose target lies in it or behind it come from the next block's snapshot
	const int H = P.comp_n ? comp_first_inverse(blk, P.n_blk_res, P.rows_in_blk[blk], BS, n_grp, P.two_sided) : n_grp;

	// ---- this thread's pairs: target row, group of the column, start rank; pairs outside the scanned rows are dropped
	uint32_t r[C], tgt[C], grp[C], valid = 0, back = 0;
	int f_max = 0, b_min = n_grp;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const int e = slice_base + warp * (C * 32) + lane * C + c;   // a lane's C pairs are neighbours in the sorted list (same rows in phase B), a warp's 32*C too
		tgt[c] = 0xffffffffu; r[c] = 0; grp[c] = 0;
		if (e < n_pairs) {
			const uint32_t t = qrow[e];
			const long long arow = blk_row + t;
			if (arow >= P.row_lo && arow < P.row_hi) {
				const int32_t col = qcol[e];
				const int tg = (int)(t / COMP_K);
				const bool bk = tg >= H;
				tgt[c] = t; valid |= 1u << c; back |= (bk ? 1u : 0u) << c;
				grp[c] = P.tgrp[col];
				r[c] = (uint32_t)P.rank0[((size_t)(blk + (bk ? 1 : 0)) * 2 + 0) * m + col];   // pbwt.c:343 under this / the next snapshot
				if (bk) b_min = min(b_min, tg); else f_max = max(f_max, tg);
			}
		}
	}
	// the groups this CTA crosses by composite map: the forward maps in front of its last forward target's group and the inverse
	// maps behind its first backward target's group, as far as maps are available
	f_max = __reduce_max_sync(PW_FULL, f_max);
	b_min = __reduce_min_sync(PW_FULL, b_min);
	if (tid == 0) { S.f_max = 0; S.b_min = n_grp; for (int s = 0; s < PW_STAGES; ++s) pw_mbar_init(&S.bar[s], 1); }
	__syncthreads();
	if (lane == 0) { atomicMax(&S.f_max, f_max); atomicMin(&S.b_min, b_min); }
	__syncthreads();
	f_max = S.f_max; b_min = S.b_min;
	if (tid == 0) { S.f_avail = f_max; S.b_avail = b_min; }
	__syncthreads();
	for (int i = tid; i < n_grp; i += PW_NT) {
		const bool have = P.comp_n != nullptr && P.comp_n[(size_t)blk * n_grp + i] != 0;
		if (!have && i < f_max) atomicMin(&S.f_avail, i);
		if (!have && i > b_min) atomicMax(&S.b_avail, i);
	}
	__syncthreads();
	const int gF = S.f_avail;                                   // forward maps of groups [0, gF) are crossed
	const int gB = S.b_avail;                                   // inverse maps of groups (gB, n_grp) are crossed, last one first
	const int nB = b_min < n_grp ? n_grp - 1 - gB : 0, nA = gF + nB;

	const long long t_a = P.prof ? clock64() : 0;
	// ---- phase A: composite maps through three TMA stages; step s crosses group s (s < gF) or group n_grp-1-(s-gF)
	{
		const uint32_t stg0 = pw_smem_u32(pw_sm);
		auto group_of = [&](int s2) { return s2 < gF ? s2 : n_grp - 1 - (s2 - gF); };
		auto fetch = [&](int s2) {
			const size_t slot = (size_t)blk * n_grp + group_of(s2);
			const uint32_t np = (uint32_t)P.comp_n[slot];
			uint64_t *bar = &S.bar[s2 % PW_STAGES];
			uint8_t *dst = pw_sm + (size_t)(s2 % PW_STAGES) * PW_STG_BYTES;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			pw_mbar_expect_tx(bar, np * 8u + (uint32_t)P.dir_n * 2u);
			pw_tma_g2s(dst, P.comp_start + slot * COMP_CAP, np * 4u, bar);
			pw_tma_g2s(dst + PW_STG_TD, P.comp_delta + slot * COMP_CAP, np * 4u, bar);
			pw_tma_g2s(dst + PW_STG_DIR, P.comp_dir + slot * COMP_DIR_STRIDE, (uint32_t)P.dir_n * 2u, bar);
		};
		uint32_t tgq[C];                                        // target group, or a value that keeps the pair out of every step
		#pragma unroll
		for (int c = 0; c < C; ++c) tgq[c] = tgt[c] / COMP_K;
		if (tid == 0) for (int k = 0; k < PW_STAGES - 1 && k < nA; ++k) fetch(k);
		for (int s2 = 0; s2 < nA; ++s2) {
			if (tid == 0 && s2 + PW_STAGES - 1 < nA) fetch(s2 + PW_STAGES - 1);   // that stage was released by the barrier that ended step s2-1
			{
				uint32_t spins = 0;
				const uint32_t par = (uint32_t)(s2 / PW_STAGES) & 1u;
				while (!pw_mbar_try_wait(&S.bar[s2 % PW_STAGES], par))
					if (++spins > (1u << 26)) { atomicOr(P.err, 8); __trap(); }
			}
			const bool fw = s2 < gF;
			const uint32_t g = (uint32_t)group_of(s2);
			uint32_t act = 0;
			#pragma unroll
			for (int c = 0; c < C; ++c) act |= (fw ? tgq[c] > g && !((back >> c) & 1u) : tgq[c] < g && ((back >> c) & 1u)) ? 1u << c : 0u;
			act &= valid;
			const uint32_t tab = stg0 + (uint32_t)(s2 % PW_STAGES) * PW_STG_BYTES;
			pw_lookup_comp<C>(r, tab, tab + PW_STG_DIR, P.dir_shift, act);
			__syncthreads();
		}
	}

	// ---- phase B: every warp on its own.  A forward pair walks from the start of its group (or of the first group without a map)
	// to its target row, whose run gives the bit; a backward pair is undone from the end of its group (or of the last group
	// without a map) down to the row behind its target, where its rank tells the bit.  The 32*C pairs of a warp are neighbours
	// in the sorted list: a few rows apart.
	const long long t_b = P.prof ? clock64() : 0;
	unsigned rows_walked = 0;
	uint8_t *slice = pw_sm + (size_t)warp * PW_SLICE;
	uint32_t *ts = (uint32_t*)slice;
	uint32_t *raw32 = (uint32_t*)(slice + PW_TAB2 * 8);
	const uint8_t *raw = (const uint8_t*)raw32;
	const int per_row = P.G * 3;
	// pairs whose target is `row`: code 3 (other-ALT) if the plane-0 bit is set, else 2 (missing) -- bgt.c:743-756
	auto count_hits = [&](uint32_t row, uint32_t hit, uint32_t bits) {
		#pragma unroll
		for (int c = 0; c < C; ++c)
			if ((hit >> c) & 1u)
				atomicAdd(P.cnt_raw + (size_t)(blk_row + row - P.row_lo) * per_row + grp[c] * 3 + (((bits >> c) & 1u) ? 2 : 1), 1);
	};
	const uint32_t fwd = valid & ~back;
	if (__any_sync(PW_FULL, fwd != 0)) { // ---- forward pairs, rows ascending
		uint32_t from[C], rlo = 0xffffffffu, rhi = 0;
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const bool v = (fwd >> c) & 1u;
			const uint32_t tg = tgt[c] / COMP_K;
			from[c] = v ? (tg < (uint32_t)gF ? tg : (uint32_t)gF) * COMP_K : 0xffffffffu;
			rlo = min(rlo, from[c]);
			rhi = max(rhi, v ? tgt[c] : 0u);
		}
		rlo = __reduce_min_sync(PW_FULL, rlo);
		rhi = __reduce_max_sync(PW_FULL, rhi);
		auto live = [&](uint32_t row) {
			uint32_t a = 0;
			#pragma unroll
			for (int c = 0; c < C; ++c) a |= (row >= from[c] && row <= tgt[c] ? 1u : 0u) << c;
			return a & fwd;
		};
		auto hits = [&](uint32_t row, uint32_t act) {
			uint32_t h = 0;
			#pragma unroll
			for (int c = 0; c < C; ++c) h |= (row == tgt[c] ? 1u : 0u) << c;
			return h & act;
		};
		for (uint32_t row = rlo; row <= rhi;) {
			// the records of up to 32 rows lie back to back in the image: stage as many whole rows as fit the slice with one
			// round of independent coalesced loads, then work on them out of shared memory
			const uint32_t nmax = rhi - row + 1 < 32u ? rhi - row + 1 : 32u;
			const uint32_t j = (uint32_t)lane < nmax ? (uint32_t)lane : nmax - 1;
			const uint64_t off_l = roff[row + j], end_l = roff[row + j + 1];
			const uint32_t n1_l = n1p[(size_t)(row + j) * 2];
			const uint64_t base = __shfl_sync(PW_FULL, off_l, 0) & ~(uint64_t)3;
			const uint32_t nb = __popc(__ballot_sync(PW_FULL, (uint32_t)lane < nmax && end_l - base <= (uint64_t)PW_RAW));
			if (nb == 0) { // a single record larger than the slice: its bytes come straight from global memory
				const uint32_t act = live(row);
				const uint32_t n1 = __shfl_sync(PW_FULL, n1_l, 0);
				uint32_t bits = n1 ? 0xffffffffu : 0u;
				if (n1 != 0 && n1 != m && __any_sync(PW_FULL, act != 0)) {
					const uint8_t *rec = P.img + __shfl_sync(PW_FULL, off_l, 0);
					bits = 0;
					pw_row<C, false>(rec + 5, pw_ld_u32_unaligned(rec + 1), m, n1, ts, lane, r, act, bits);
				}
				count_hits(row, hits(row, act), bits);
				++rows_walked; ++row;
				continue;
			}
			const uint32_t words = (uint32_t)((__shfl_sync(PW_FULL, end_l, nb - 1) - base + 3) >> 2);
			const uint32_t *src = (const uint32_t*)(P.img + base);
			__syncwarp();
			for (uint32_t w = lane; w < words; w += 32) raw32[w] = src[w];
			__syncwarp();
			for (uint32_t k = 0; k < nb; ++k, ++row) {
				const uint32_t act = live(row);
				if (!__any_sync(PW_FULL, act != 0)) continue;
				++rows_walked;
				const uint32_t n1 = __shfl_sync(PW_FULL, n1_l, k);
				uint32_t bits = n1 ? 0xffffffffu : 0u;         // constant row: the order does not change (pbwt.c:75-77)
				if (n1 != 0 && n1 != m) {
					const uint8_t *rec = raw + (uint32_t)(__shfl_sync(PW_FULL, off_l, k) - base);
					const uint32_t l0 = (uint32_t)rec[1] | (uint32_t)rec[2] << 8 | (uint32_t)rec[3] << 16 | (uint32_t)rec[4] << 24;
					bits = 0;
					pw_row<C, false>(rec + 5, l0, m, n1, ts, lane, r, act, bits);
				}
				count_hits(row, hits(row, act), bits);
			}
		}
	}
	if (__any_sync(PW_FULL, back != 0)) { // ---- backward pairs, rows descending
		uint32_t top[C], rlo = 0xffffffffu, rhi = 0;            // top[c]: first row to undo (the last row of the group the maps stop at)
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const bool v = (back >> c) & 1u;
			const uint32_t tg = tgt[c] / COMP_K;
			top[c] = v ? min(((tg > (uint32_t)gB ? tg : (uint32_t)gB) + 1u) * COMP_K, (uint32_t)BS) - 1u : 0u;   // (a block with inverse maps is full: BS rows)
			rlo = min(rlo, v ? tgt[c] : 0xffffffffu);
			rhi = max(rhi, top[c]);
		}
		rlo = __reduce_min_sync(PW_FULL, rlo);
		rhi = __reduce_max_sync(PW_FULL, rhi);
		for (uint32_t row = rhi; (int)row >= (int)rlo;) {
			// rows [row - nmax + 1, row]: stage the longest tail of them that fits, work downwards
			const uint32_t nmax = row - rlo + 1 < 32u ? row - rlo + 1 : 32u;
			const uint32_t lo = row - nmax + 1;
			const uint32_t j = (uint32_t)lane < nmax ? (uint32_t)lane : nmax - 1;
			const uint64_t off_l = roff[lo + j];
			const uint32_t n1_l = n1p[(size_t)(lo + j) * 2];
			const uint64_t end_top = roff[row + 1];
			const uint32_t nb = __popc(__ballot_sync(PW_FULL, (uint32_t)lane < nmax && end_top - (off_l & ~(uint64_t)3) <= (uint64_t)PW_RAW));
			const bool staged = nb > 0;
			const uint32_t first = staged ? nmax - nb : nmax - 1; // lane that holds the lowest row handled in this round
			const uint64_t base = __shfl_sync(PW_FULL, off_l, first) & ~(uint64_t)3;
			if (staged) {
				const uint32_t words = (uint32_t)((end_top - base + 3) >> 2);
				const uint32_t *src = (const uint32_t*)(P.img + base);
				__syncwarp();
				for (uint32_t w = lane; w < words; w += 32) raw32[w] = src[w];
				__syncwarp();
			}
			for (uint32_t k = nmax; k-- > first; --row) {
				uint32_t undo = 0, hit = 0;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					undo |= (row > tgt[c] && row <= top[c] ? 1u : 0u) << c;
					hit |= (row == tgt[c] ? 1u : 0u) << c;
				}
				undo &= back; hit &= back;
				if (!__any_sync(PW_FULL, (undo | hit) != 0)) continue;
				++rows_walked;
				const uint32_t n1 = __shfl_sync(PW_FULL, n1_l, k);
				if (n1 != 0 && n1 != m && __any_sync(PW_FULL, undo != 0)) {
					const uint64_t off = __shfl_sync(PW_FULL, off_l, k);
					const uint8_t *rec = staged ? raw + (uint32_t)(off - base) : P.img + off;
					const uint32_t l0 = staged ? ((uint32_t)rec[1] | (uint32_t)rec[2] << 8 | (uint32_t)rec[3] << 16 | (uint32_t)rec[4] << 24) : pw_ld_u32_unaligned(rec + 1);
					uint32_t unused = 0;
					pw_row<C, true>(rec + 5, l0, m, n1, ts, lane, r, undo, unused);
				}
				if (__any_sync(PW_FULL, hit != 0)) { // rank behind the target row: the ones sit behind the m - n1 zeros (pbwt.c:79-88)
					uint32_t bits = 0;
					#pragma unroll
					for (int c = 0; c < C; ++c) bits |= (r[c] >= m - n1 ? 1u : 0u) << c;
					count_hits(row, hit, bits);
				}
			}
		}
	}
	if (P.prof) {
		if (lane == 0) atomicAdd(P.prof + 4, (unsigned long long)rows_walked);
		__syncthreads();
		if (tid == 0) {
			const long long t_e = clock64();
			atomicAdd(P.prof + 0, (unsigned long long)(t_b - t_a)); atomicAdd(P.prof + 1, (unsigned long long)(t_e - t_b));
			atomicAdd(P.prof + 2, 1ull); atomicAdd(P.prof + 3, (unsigned long long)nA);
		}
	}
}

size_t pair_smem_bytes() { return (size_t)PW_STAGES * PW_STG_BYTES; }

template<int C>
static cudaError_t launch_pair_t(const PairParams &P, int max_pairs, int n_blk, cudaStream_t st)
{
	const size_t smem = pair_smem_bytes();
	cudaError_t e = cudaFuncSetAttribute(pbwt_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	const int slices = (max_pairs + PW_NT * C - 1) / (PW_NT * C);
	for (int b0 = 0; b0 < n_blk; b0 += 32768) {
		PairParams Q = P;
		if (P.blk_list) Q.blk_list = P.blk_list + b0; else Q.blk_first = P.blk_first + b0;
		const int nb = n_blk - b0 < 32768 ? n_blk - b0 : 32768;
		pbwt_pair_kernel<C><<<dim3(slices, nb), PW_NT, smem, st>>>(Q);
	}
	return cudaGetLastError();
}

cudaError_t launch_pairwalk(const PairParams &P, int C, int max_pairs, int n_blk, cudaStream_t st)
{
	if (max_pairs <= 0 || n_blk <= 0) return cudaSuccess;
	switch (C) {
	case 1: return launch_pair_t<1>(P, max_pairs, n_blk, st);
	case 4: return launch_pair_t<4>(P, max_pairs, n_blk, st);
	default: return launch_pair_t<2>(P, max_pairs, n_blk, st);
	}
}

} // namespace b200
