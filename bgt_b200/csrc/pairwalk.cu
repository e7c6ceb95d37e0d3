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

namespace b200 {

#define PW_FULL 0xffffffffu

constexpr int PW_NT = 512, PW_NW = PW_NT / 32;
constexpr int PW_STAGES = 3;
constexpr uint32_t PW_STG_TD = COMP_CAP * 4u, PW_STG_DIR = COMP_CAP * 8u, PW_STG_BYTES = COMP_CAP * 8u + COMP_DIR_STRIDE * 2u;
constexpr int PW_TAB = (int)((PW_STAGES * PW_STG_BYTES) / PW_NW / 8) & ~31;   // run-table entries per warp in phase B (overlays the stages)

__device__ __forceinline__ uint32_t pw_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pw_lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t pw_lds_u16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t pw_rle_len(uint32_t c) { const uint32_t v = c >> 1; return (v & 15u) << ((v >> 4) << 2); }

__device__ __forceinline__ uint32_t pw_ld_u32_unaligned(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	const uint32_t lo = w[0];
	if (sh == 0) return lo;
	return __funnelshift_r(lo, w[1], sh);
}

__device__ __forceinline__ void pw_mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(pw_smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pw_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pw_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool pw_mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(pw_smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void pw_tma_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(pw_smem_u32(dst)), "l"(src), "r"(bytes), "r"(pw_smem_u32(bar)) : "memory");
}

// one look-up per slot in a staged composite map through its bucket directory: the piece of rank r lies between dir[r >> s]
// and dir[(r >> s) + 1]; the widest window of the warp sets the trip count.  Only the slots in `act` move.
template<int C>
__device__ __forceinline__ void pw_lookup_comp(uint32_t (&r)[C], uint32_t tab, uint32_t dir, int sh, uint32_t act)
{
	if (!__any_sync(PW_FULL, act != 0)) return;
	uint32_t a[C], hi[C], w = 1;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t d = dir + ((r[c] >> sh) << 1);
		const uint32_t lo = pw_lds_u16(d), h = pw_lds_u16(d + 2);
		a[c] = tab + (lo << 2); hi[c] = tab + (h << 2);
		const uint32_t wc = ((act >> c) & 1u) ? h - lo + 1u : 1u;
		w = wc > w ? wc : w;
	}
	w = __reduce_max_sync(PW_FULL, w);
	for (uint32_t len = w; len > 1;) {
		const uint32_t half = len >> 1, h4 = half << 2;
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const uint32_t t = a[c] + h4;
			const uint32_t v = t <= hi[c] ? pw_lds_u32(t) : 0xffffffffu;
			a[c] = v <= r[c] ? t : a[c];
		}
		len -= half;
	}
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t d = pw_lds_u32(a[c] + PW_STG_TD);
		if ((act >> c) & 1u) r[c] += d;
	}
}

// One warp, one row, forward: the lanes in `act` move their rank through the row (rank' = rank + delta of the run that
// holds it, pbwt.c:150) and learn the run's bit.  The row's RLE bytes come straight from global memory; its run table is
// built piecewise in the warp's shared-memory slice (ts = run starts, td = deltas, PW_TAB entries a piece).
__device__ __forceinline__ void pw_row_forward(const uint8_t *rle, uint32_t l, uint32_t m, uint32_t n1, uint32_t *ts, int32_t *td, int lane,
                                               uint32_t &r, bool act, uint32_t &bit)
{
	const uint32_t zeros_total = m - n1;
	uint32_t tot = 0, ones = 0;
	bool done = !act;
	for (uint32_t cb = 0; cb < l; cb += PW_TAB) {
		const uint32_t n = l - cb < (uint32_t)PW_TAB ? l - cb : (uint32_t)PW_TAB;
		const uint32_t cs = tot;
		for (uint32_t base = 0; base < n; base += 32) {
			const uint32_t i = base + lane;
			const uint32_t c = i < n ? rle[cb + i] : 0u;
			const uint32_t L = pw_rle_len(c), b = c & 1u, L1 = b ? L : 0u;
			uint32_t x = L, y = L1;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t tx = __shfl_up_sync(PW_FULL, x, d), ty = __shfl_up_sync(PW_FULL, y, d);
				if (lane >= d) { x += tx; y += ty; }
			}
			const uint32_t start = tot + x - L, ones_before = ones + y - L1;
			if (i < n) { ts[i] = start; td[i] = b ? (int32_t)(zeros_total - (start - ones_before)) : -(int32_t)ones_before; }
			tot += __shfl_sync(PW_FULL, x, 31);
			ones += __shfl_sync(PW_FULL, y, 31);
		}
		__syncwarp();
		if (!done && r >= cs && r < tot) { // last entry whose start <= r (zero-length bytes share the start of their successor and are never the last)
			uint32_t a = 0;
			for (uint32_t len = n; len > 1;) { const uint32_t half = len >> 1; a += ts[a + half] <= r ? half : 0; len -= half; }
			r += (uint32_t)td[a];
			bit = r >= zeros_total ? 1u : 0u;
			done = true;
		}
		__syncwarp();
		if (__all_sync(PW_FULL, done)) break;
	}
}

struct PairSmem { uint64_t bar[PW_STAGES]; int avail; };

template<int C>
__global__ void __launch_bounds__(PW_NT, 2) pbwt_pair_kernel(const PairParams P)
{
	extern __shared__ __align__(128) uint8_t pw_sm[];
	__shared__ PairSmem S;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int BS = 1 << P.shift;
	const int blk = P.blk_list ? P.blk_list[blockIdx.y] : P.blk_first + (int)blockIdx.y;
	const int n_pairs = P.qcount[blk];
	const int slice_base = blockIdx.x * (PW_NT * C);
	if (slice_base >= n_pairs) return;
	const uint32_t m = (uint32_t)P.m;
	const int32_t *qcol = P.qcol + (size_t)blk * P.q_stride;
	const uint16_t *qrow = P.qrow + (size_t)blk * P.q_stride;
	const long long blk_row = P.blk_row0 + ((long long)blk << P.shift);
	const int n_grp = (BS + COMP_K - 1) / COMP_K;
	const uint64_t *roff = P.rowoff + (size_t)blk * (BS + 1);
	const uint32_t *n1p = P.n1 + (size_t)blk * BS * 2;

	// ---- this thread's pairs: target row, group of the column, start rank; pairs outside the scanned rows are dropped
	uint32_t r[C], tgt[C], grp[C], valid = 0;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const int e = slice_base + c * PW_NT + tid;
		tgt[c] = 0xffffffffu; r[c] = 0; grp[c] = 0;
		if (e < n_pairs) {
			const uint32_t t = qrow[e];
			const long long arow = blk_row + t;
			if (arow >= P.row_lo && arow < P.row_hi) {
				const int32_t col = qcol[e];
				tgt[c] = t; valid |= 1u << c;
				grp[c] = P.tgrp[col];
				r[c] = (uint32_t)P.rank0[((size_t)blk * 2 + 0) * m + col];
			}
		}
	}
	// the groups this CTA crosses by composite: those in front of its last live target's group, as far as maps are available
	int g_last = 0;
	#pragma unroll
	for (int c = 0; c < C; ++c) if ((valid >> c) & 1u) g_last = max(g_last, (int)(tgt[c] / COMP_K));
	g_last = __reduce_max_sync(PW_FULL, g_last);
	if (tid == 0) { S.avail = 0; for (int s = 0; s < PW_STAGES; ++s) pw_mbar_init(&S.bar[s], 1); }
	__syncthreads();
	if (lane == 0) atomicMax(&S.avail, g_last);
	__syncthreads();
	g_last = S.avail;
	__syncthreads();
	if (tid == 0) S.avail = g_last;
	__syncthreads();
	for (int i = tid; i < g_last; i += PW_NT) if (P.comp_n == nullptr || P.comp_n[(size_t)blk * n_grp + i] == 0) atomicMin(&S.avail, i);
	__syncthreads();
	const int g_end = S.avail;

	// ---- phase A: composite maps through three TMA stages
	{
		const uint32_t stg0 = pw_smem_u32(pw_sm);
		auto fetch = [&](int g) {
			const size_t slot = (size_t)blk * n_grp + g;
			const uint32_t np = (uint32_t)P.comp_n[slot];
			uint64_t *bar = &S.bar[g % PW_STAGES];
			uint8_t *dst = pw_sm + (size_t)(g % PW_STAGES) * PW_STG_BYTES;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			pw_mbar_expect_tx(bar, np * 8u + (uint32_t)P.dir_n * 2u);
			pw_tma_g2s(dst, P.comp_start + slot * COMP_CAP, np * 4u, bar);
			pw_tma_g2s(dst + PW_STG_TD, P.comp_delta + slot * COMP_CAP, np * 4u, bar);
			pw_tma_g2s(dst + PW_STG_DIR, P.comp_dir + slot * COMP_DIR_STRIDE, (uint32_t)P.dir_n * 2u, bar);
		};
		if (tid == 0) for (int k = 0; k < PW_STAGES - 1 && k < g_end; ++k) fetch(k);
		for (int g = 0; g < g_end; ++g) {
			if (tid == 0 && g + PW_STAGES - 1 < g_end) fetch(g + PW_STAGES - 1);   // that stage was released by the barrier that ended group g-1
			{
				uint32_t spins = 0;
				const uint32_t par = (uint32_t)(g / PW_STAGES) & 1u;
				while (!pw_mbar_try_wait(&S.bar[g % PW_STAGES], par))
					if (++spins > (1u << 26)) { atomicOr(P.err, 8); __trap(); }
			}
			uint32_t act = 0;
			#pragma unroll
			for (int c = 0; c < C; ++c) act |= (((valid >> c) & 1u) && (int)(tgt[c] / COMP_K) > g ? 1u : 0u) << c;
			const uint32_t tab = stg0 + (uint32_t)(g % PW_STAGES) * PW_STG_BYTES;
			pw_lookup_comp<C>(r, tab, tab + PW_STG_DIR, P.dir_shift, act);
			__syncthreads();
		}
	}

	// ---- phase B: every warp on its own, slot by slot.  A pair walks from the start of its group (or of the first group
	// without a composite map) to its target row; the 32 pairs of a slot are neighbours in the sorted list.
	uint32_t *ts = (uint32_t*)pw_sm + (size_t)warp * PW_TAB * 2;
	int32_t *td = (int32_t*)(ts + PW_TAB);
	const int per_row = P.G * 3;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const bool v = (valid >> c) & 1u;
		const uint32_t tg = tgt[c] / COMP_K;
		const uint32_t from = v ? (tg < (uint32_t)g_end ? tg : (uint32_t)g_end) * COMP_K : 0xffffffffu;
		const uint32_t rlo = __reduce_min_sync(PW_FULL, from);
		const uint32_t rhi = __reduce_max_sync(PW_FULL, v ? tgt[c] : 0u);
		if (rlo == 0xffffffffu) continue;                      // no live pair in this slot of the warp
		for (uint32_t row = rlo; row <= rhi; ++row) {
			const bool act = v && row >= from && row <= tgt[c];
			if (!__any_sync(PW_FULL, act)) continue;
			const uint32_t n1 = n1p[(size_t)row * 2];
			uint32_t bit = n1 ? 1u : 0u;                       // constant row: the order does not change (pbwt.c:75-77)
			if (n1 != 0 && n1 != m) {
				const uint8_t *rec = P.img + roff[row];
				const uint32_t l0 = pw_ld_u32_unaligned(rec + 1);
				pw_row_forward(rec + 5, l0, m, n1, ts, td, lane, r[c], act, bit);
			}
			// pairs whose target is this row: code 3 (other-ALT) if the plane-0 bit is set, else 2 (missing) -- bgt.c:743-756
			const bool hit = act && row == tgt[c];
			if (__any_sync(PW_FULL, hit)) {
				const uint32_t key = hit ? (grp[c] << 1 | bit) : 0xffffffffu;
				const uint32_t peers = __match_any_sync(PW_FULL, key);
				if (hit && lane == __ffs(peers) - 1)
					atomicAdd(P.cnt_raw + (size_t)(blk_row + row - P.row_lo) * per_row + grp[c] * 3 + (bit ? 2 : 1), __popc(peers));
			}
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
