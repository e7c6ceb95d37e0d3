// plane1.cu -- first phase of the split scan: WHICH haplotypes carry a plane-1 bit (missing / other-ALT code), and
// at which rows, without walking all m columns.
//
// The second bit plane is empty on most rows and sparse on the rest, so instead of advancing the rank of every
// column (m look-ups per non-empty row) the question is asked backwards: the t-th 1 bit of row v sits at a known rank of
// the PBWT order in front of row v; the non-empty plane-1 rows before it in the block are undone (the inverse of the
// stable partition of pbwt.c:79-88: a rank in the zeros part came from the k-th 0 of the row, a rank in the ones
// part from the k-th 1) until the block's snapshot is reached, where S[rank] (pbwt.c:298-300) names the column.
// Cost: (ones in the block's plane 1) x (non-empty rows before them, crossed 32 at a time) -- independent of m.
//
// The block's non-empty plane-1 rows are its "view" (index.cu: p1view_kernel): re-framed records back to back, their
// ones counted in front of every view row (prefix), inverse composite maps of every 32 view rows (compose.cu).  The kernel
// has the shape of the pair walk (pairwalk.cu) run the other way round: grid = (slices of the block's plane-1 bits in
// row order, blocks), 512 threads, C bits per thread;
//   rows   first (warp-autonomous): a bit starts BEHIND its own view row at rank zeros + t -- where the partition puts the
//          t-th 1 -- and the view rows from its own down to the start of its 32-row group are undone one by one (run tables
//          the warp builds for itself from staged record bytes);
//   groups then (CTA-wide): the groups in front of it are crossed, last one first, by inverse composite maps streamed
//          through shared memory (TMA bulk copies into three stages, mbarrier completion), one look-up each through the map's
//          bucket directory;
//   then   S[rank] under the block's plane-1 snapshot is the column.
// Output per block, in row order: qcol[] / qrow[] = (column, row within the block) of every plane-1 bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"
#include "pairwalk.cuh"

namespace b200 {

constexpr uint32_t P1_STG_TD = SELECT_COMP_CAP * 4u, P1_STG_DIR = SELECT_COMP_CAP * 8u, P1_STG_BYTES = SELECT_COMP_CAP * 8u + COMP_DIR_STRIDE * 2u;
constexpr int P1_SLICE = (int)((PW_STAGES * P1_STG_BYTES) / PW_NW) & ~15;     // a warp's share of the (idle) stages while rows are undone
constexpr int P1_RAW = P1_SLICE - PW_TAB2 * 8;                                // staged record bytes per warp

struct SelSmem { uint64_t bar[PW_STAGES]; int g_max, g_un; };

template<int C>
__global__ void __launch_bounds__(PW_NT, 2) plane1_select_kernel(const SelectParams P)
{
	extern __shared__ __align__(128) uint8_t p1_sm[];
	__shared__ SelSmem S;
	const int blk = P.blk_list ? P.blk_list[blockIdx.y] : P.blk_first + (int)blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (P.blk_ok && !P.blk_ok[blk]) { if (tid == 0 && blockIdx.x == 0 && P.slice0 == 0) P.qcount[blk] = 0; return; }   // not a sparse block: the general walk takes it
	const int nv = P.p1_rows_in_blk[blk];
	const long long vb = P.p1_vbase[blk];
	const uint64_t *ro = P.p1_rowoff + vb + blk;
	const uint32_t *prefix = P.p1_prefix + vb + blk;
	const uint32_t *n1v = P.p1_n1 + vb;
	const uint32_t Q = prefix[nv];
	if (Q > (uint32_t)P.cap) { if (tid == 0 && blockIdx.x == 0 && P.slice0 == 0) { atomicOr(P.err, 32); P.qcount[blk] = 0; } return; }
	if (tid == 0 && blockIdx.x == 0 && P.slice0 == 0) P.qcount[blk] = (int)Q;
	const uint32_t slice_base = (blockIdx.x + (uint32_t)P.slice0) * (uint32_t)(PW_NT * C);
	if (slice_base >= Q) return;
	const uint32_t m = (uint32_t)P.m;

	// ---- this thread's bits: view row, rank behind that row (the partition puts the t-th 1 of a row at zeros + t, pbwt.c:79-88)
	uint32_t r[C], vrow[C], valid = 0;
	int g_max = 0;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t q = slice_base + warp * (C * 32) + lane * C + c;   // a lane's C bits are neighbours in row order, a warp's 32*C too
		vrow[c] = 0; r[c] = 0;
		if (q < Q) {
			uint32_t lo = 0;                                   // last view row whose prefix <= q
			for (uint32_t len = (uint32_t)nv; len > 1;) { const uint32_t half = len >> 1; lo += prefix[lo + half] <= q ? half : 0u; len -= half; }
			vrow[c] = lo; valid |= 1u << c;
			r[c] = (m - n1v[lo]) + (q - prefix[lo]);
			g_max = max(g_max, (int)(lo / COMP_K));
		}
	}
	// Groups [g_lo, own group) are crossed by inverse composite map.  g_lo = 0 unless a map below the CTA's highest group is
	// missing (more pieces than a slot holds): then the rows in front of group g_lo are undone one by one at the end.
	g_max = __reduce_max_sync(PW_FULL, g_max);
	if (tid == 0) { S.g_max = 0; S.g_un = -1; for (int s = 0; s < PW_STAGES; ++s) pw_mbar_init(&S.bar[s], 1); }
	__syncthreads();
	if (lane == 0) atomicMax(&S.g_max, g_max);
	__syncthreads();
	g_max = S.g_max;
	for (int i = tid; i < g_max; i += PW_NT) if (P.vcomp_n == nullptr || P.vcomp_n[(size_t)blk * SELECT_GROUPS + i] == 0) atomicMax(&S.g_un, i);
	__syncthreads();
	const int g_lo = S.g_un + 1;

	uint8_t *slice = p1_sm + (size_t)warp * P1_SLICE;
	uint32_t *ts = (uint32_t*)slice;
	uint32_t *raw32 = (uint32_t*)(slice + PW_TAB2 * 8);
	const uint8_t *raw = (const uint8_t*)raw32;
	// undo the view rows [low[c], high[c]] of every live slot, last row first (warp-autonomous; rows staged in batches)
	auto undo_rows = [&](const uint32_t (&high)[C], const uint32_t (&low)[C], uint32_t live) {
		uint32_t rlo = 0xffffffffu, rhi = 0;
		#pragma unroll
		for (int c = 0; c < C; ++c) if ((live >> c) & 1u) { rlo = min(rlo, low[c]); rhi = max(rhi, high[c]); }
		rlo = __reduce_min_sync(PW_FULL, rlo);
		rhi = __reduce_max_sync(PW_FULL, rhi);
		if (rlo == 0xffffffffu) return;
		for (uint32_t row = rhi; (int)row >= (int)rlo;) {
			const uint32_t nmax = row - rlo + 1 < 32u ? row - rlo + 1 : 32u;
			const uint32_t lo = row - nmax + 1;
			const uint32_t j = (uint32_t)lane < nmax ? (uint32_t)lane : nmax - 1;
			const uint64_t off_l = ro[lo + j];
			const uint32_t n1_l = n1v[lo + j];
			const uint64_t end_top = ro[row + 1];
			const uint32_t nb = __popc(__ballot_sync(PW_FULL, (uint32_t)lane < nmax && end_top - (off_l & ~(uint64_t)3) <= (uint64_t)P1_RAW));
			const bool staged = nb > 0;
			const uint32_t first = staged ? nmax - nb : nmax - 1;
			const uint64_t base = __shfl_sync(PW_FULL, off_l, first) & ~(uint64_t)3;
			if (staged) {
				const uint32_t words = (uint32_t)((end_top - base + 3) >> 2);
				const uint32_t *src = (const uint32_t*)(P.p1img + base);
				__syncwarp();
				for (uint32_t w = lane; w < words; w += 32) raw32[w] = src[w];
				__syncwarp();
			}
			for (uint32_t k = nmax; k-- > first; --row) {
				uint32_t act = 0;
				#pragma unroll
				for (int c = 0; c < C; ++c) act |= (row >= low[c] && row <= high[c] ? 1u : 0u) << c;
				act &= live;
				if (!__any_sync(PW_FULL, act != 0)) continue;
				const uint32_t n1 = __shfl_sync(PW_FULL, n1_l, k);
				if (n1 == 0 || n1 == m) continue;                  // constant row: the order does not change
				const uint64_t off = __shfl_sync(PW_FULL, off_l, k);
				const uint8_t *rec = staged ? raw + (uint32_t)(off - base) : P.p1img + off;   // view record: 'B', l0 = 0, l1, bytes
				const uint32_t l1 = staged ? ((uint32_t)rec[5] | (uint32_t)rec[6] << 8 | (uint32_t)rec[7] << 16 | (uint32_t)rec[8] << 24) : pw_ld_u32_unaligned(rec + 5);
				uint32_t unused = 0;
				pw_row<C, true>(rec + 9, l1, m, n1, ts, lane, r, act, unused);
			}
		}
	};

	// ---- rows of the own group (of everything in front, where no map can be used)
	uint32_t high[C], low[C];
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t g = vrow[c] / COMP_K;
		high[c] = vrow[c];
		low[c] = (int)g >= g_lo ? g * COMP_K : 0u;
	}
	undo_rows(high, low, valid);
	__syncthreads();                                            // the warps' slices are the stages of the next phase

	// ---- groups in front, last one first: step s crosses group g_max - 1 - s
	{
		const uint32_t stg0 = pw_smem_u32(p1_sm);
		const int nA = g_max > g_lo ? g_max - g_lo : 0;
		auto fetch = [&](int s2) {
			const size_t slot = (size_t)blk * SELECT_GROUPS + (g_max - 1 - s2);
			const uint32_t np = (uint32_t)P.vcomp_n[slot];
			uint64_t *bar = &S.bar[s2 % PW_STAGES];
			uint8_t *dst = p1_sm + (size_t)(s2 % PW_STAGES) * P1_STG_BYTES;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			pw_mbar_expect_tx(bar, np * 8u + (uint32_t)P.dir_n * 2u);
			pw_tma_g2s(dst, P.vcomp_start + slot * SELECT_COMP_CAP, np * 4u, bar);
			pw_tma_g2s(dst + P1_STG_TD, P.vcomp_delta + slot * SELECT_COMP_CAP, np * 4u, bar);
			pw_tma_g2s(dst + P1_STG_DIR, P.vcomp_dir + slot * COMP_DIR_STRIDE, (uint32_t)P.dir_n * 2u, bar);
		};
		if (tid == 0) for (int k = 0; k < PW_STAGES - 1 && k < nA; ++k) fetch(k);
		for (int s2 = 0; s2 < nA; ++s2) {
			if (tid == 0 && s2 + PW_STAGES - 1 < nA) fetch(s2 + PW_STAGES - 1);
			{
				uint32_t spins = 0;
				const uint32_t par = (uint32_t)(s2 / PW_STAGES) & 1u;
				while (!pw_mbar_try_wait(&S.bar[s2 % PW_STAGES], par))
					if (++spins > (1u << 26)) { atomicOr(P.err, 8); __trap(); }
			}
			const uint32_t g = (uint32_t)(g_max - 1 - s2);
			uint32_t act = 0;
			#pragma unroll
			for (int c = 0; c < C; ++c) act |= (vrow[c] / COMP_K > g ? 1u : 0u) << c;
			act &= valid;
			const uint32_t tab = stg0 + (uint32_t)(s2 % PW_STAGES) * P1_STG_BYTES;
			pw_lookup_comp<C, P1_STG_TD>(r, tab, tab + P1_STG_DIR, P.dir_shift, act);
			__syncthreads();
		}
	}
	// ---- a missing map below: the rows in front of group g_lo, one by one
	if (g_lo > 0) {
		uint32_t live = 0;
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const bool on = ((valid >> c) & 1u) && (int)(vrow[c] / COMP_K) >= g_lo;
			live |= (on ? 1u : 0u) << c;
			high[c] = (uint32_t)g_lo * COMP_K - 1u; low[c] = 0;
		}
		undo_rows(high, low, live);
	}
	// ---- the column under the block's plane-1 snapshot (pbwt.c:298-300)
	const uint8_t *S1 = P.img + P.blkoff[blk] + 1 + 4 * (size_t)P.m;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		if (!((valid >> c) & 1u)) continue;
		const uint32_t q = slice_base + warp * (C * 32) + lane * C + c;
		const uint32_t col = r[c] < m ? pw_ld_u32_unaligned(S1 + 4 * (size_t)r[c]) : 0xffffffffu;
		if (col >= m) { atomicOr(P.err, 64); continue; }
		const long long at = (P.ext_off ? P.ext_off[blk] - P.ext_shift : (long long)blk * (P.q_stride ? P.q_stride : (long long)P.cap)) + q;
		P.qcol[at] = (int32_t)col;
		P.qrow[at] = P.p1_realrow[vb + vrow[c]];
	}
}

cudaError_t launch_plane1_select(const SelectParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	constexpr int C = 4;
	const size_t smem = (size_t)PW_STAGES * P1_STG_BYTES;
	cudaError_t e = cudaFuncSetAttribute(plane1_select_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	const int slices = P.n_slices > 0 ? P.n_slices : (P.cap + PW_NT * C - 1) / (PW_NT * C) - P.slice0;
	if (slices <= 0) return cudaSuccess;
	for (int b0 = 0; b0 < n_blk; b0 += 32768) {
		SelectParams Q = P;
		if (P.blk_list) Q.blk_list = P.blk_list + b0; else Q.blk_first = P.blk_first + b0;
		const int nb = n_blk - b0 < 32768 ? n_blk - b0 : 32768;
		plane1_select_kernel<C><<<dim3(slices, nb), PW_NT, smem, st>>>(Q);
	}
	return cudaGetLastError();
}

} // namespace b200
