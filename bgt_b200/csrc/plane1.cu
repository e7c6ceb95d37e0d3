// plane1.cu -- first phase of the split scan: WHICH haplotypes carry a plane-1 bit (missing / other-ALT code), and
// at which rows, without walking all m columns.
//
// The second bit plane is empty on most rows and sparse on the rest, so instead of advancing the rank of every
// column (m look-ups per non-empty row) the question is asked backwards: a 1 bit of row v sits at a known rank of
// the PBWT order in front of row v; the rows before it in the block are undone one by one (the inverse of the
// stable partition of pbwt.c:79-88: a rank in the zeros part came from the k-th 0 of the row, a rank in the ones
// part from the k-th 1) until the block's snapshot is reached, where S[rank] (pbwt.c:298-300) names the column.
// Cost: (ones in the block's plane 1) x (non-empty rows before them) x (a dozen RLE bytes) -- independent of m.
//
// One CTA per checkpoint block; the block's plane-1 view (api.cu: build_plane1_view) is staged in shared memory.
// Output per block, in row order: qcol[] / qrow[] = (column, row within the block) of every plane-1 bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t p1_rle_len(uint32_t c) { const uint32_t v = c >> 1; return (v & 15u) << ((v >> 4) << 2); }

__device__ __forceinline__ uint32_t p1_ld_u32_unaligned(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	const uint32_t lo = w[0];
	if (sh == 0) return lo;
	return __funnelshift_r(lo, w[1], sh);
}

// rank (in the order in front of the row) of the t-th 1 of the row
__device__ __forceinline__ uint32_t rank_of_one(const uint8_t *rle, uint32_t l, uint32_t t)
{
	uint32_t start = 0, ones = 0;
	for (uint32_t i = 0; i < l; ++i) {
		const uint32_t c = rle[i], len = p1_rle_len(c);
		if (c & 1) { if (t < ones + len) return start + (t - ones); ones += len; }
		start += len;
	}
	return 0xffffffffu;
}

// inverse of the row's partition: rank behind the row -> rank in front of it
__device__ __forceinline__ uint32_t undo_row(const uint8_t *rle, uint32_t l, uint32_t zeros_total, uint32_t r)
{
	const uint32_t want_bit = r >= zeros_total ? 1u : 0u;
	const uint32_t k = want_bit ? r - zeros_total : r;   // k-th element of its class
	uint32_t start = 0, seen = 0;
	for (uint32_t i = 0; i < l; ++i) {
		const uint32_t c = rle[i], len = p1_rle_len(c);
		if ((c & 1u) == want_bit) { if (k < seen + len) return start + (k - seen); seen += len; }
		start += len;
	}
	return 0xffffffffu;
}

__global__ void __launch_bounds__(1024) plane1_select_kernel(const SelectParams P)
{
	extern __shared__ __align__(16) uint8_t sm[];
	uint32_t *prefix = (uint32_t*)sm;                               // [SELECT_MAX_ROWS + 1] ones before view row v
	uint32_t *roff = prefix + SELECT_MAX_ROWS + 1;                  // [SELECT_MAX_ROWS + 1] record offsets inside raw
	uint32_t *n1v = roff + SELECT_MAX_ROWS + 1;                     // [SELECT_MAX_ROWS]
	uint8_t *raw = (uint8_t*)(n1v + SELECT_MAX_ROWS);               // [SELECT_MAX_BYTES]
	uint32_t *cs = (uint32_t*)(raw + SELECT_MAX_BYTES);             // [SELECT_COMP_SMEM] inverse composites of the block's row groups, packed
	int32_t *cd = (int32_t*)(cs + SELECT_COMP_SMEM);                // [SELECT_COMP_SMEM]
	__shared__ uint32_t warp_tot[32];
	__shared__ int cg_off[SELECT_GROUPS + 1];                       // first piece of group g in cs/cd; cg_off[g+1]-cg_off[g] = 0: not available

	const int blk = P.blk_list ? P.blk_list[blockIdx.x] : P.blk_first + (int)blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (P.blk_ok && !P.blk_ok[blk]) { if (tid == 0) P.qcount[blk] = 0; return; }   // not a sparse block: the general walk takes it
	const int nv = P.p1_rows_in_blk[blk];
	const long long vb = P.p1_vbase[blk];
	const uint64_t *ro = P.p1_rowoff + vb + blk;
	const uint64_t base = ro[0];
	const uint32_t nbytes = (uint32_t)(ro[nv] - base);
	if (nv >= SELECT_MAX_ROWS || nbytes > (uint32_t)SELECT_MAX_BYTES) { if (tid == 0) { atomicOr(P.err, 16); P.qcount[blk] = 0; } return; }
	for (int v = tid; v <= nv; v += 1024) roff[v] = (uint32_t)(ro[v] - base);
	for (int v = tid; v < nv; v += 1024) n1v[v] = P.p1_n1[vb + v];
	for (uint32_t i = tid; i < nbytes; i += 1024) raw[i] = P.p1img[base + i];
	__syncthreads();
	// exclusive prefix of the per-row ones: 4 rows per thread, warp scan, cross-warp fix-up
	{
		uint32_t x[4], s = 0;
		#pragma unroll
		for (int j = 0; j < 4; ++j) { const int v = tid * 4 + j; x[j] = v < nv ? n1v[v] : 0u; s += x[j]; }
		uint32_t incl = s;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
		if (lane == 31) warp_tot[warp] = incl;
		__syncthreads();
		uint32_t before = incl - s;
		for (int w = 0; w < warp; ++w) before += warp_tot[w];
		#pragma unroll
		for (int j = 0; j < 4; ++j) { const int v = tid * 4 + j; if (v <= nv) prefix[v] = before; before += x[j]; }
	}
	__syncthreads();
	// inverse composite maps of the full COMP_K-row groups of the view (compose.cu), as many as fit
	const int n_full = nv / COMP_K;
	if (tid == 0) {
		int acc = 0;
		for (int g = 0; g < SELECT_GROUPS; ++g) {
			cg_off[g] = acc;
			const int np = (P.vcomp_n && g < n_full) ? P.vcomp_n[(size_t)blk * SELECT_GROUPS + g] : 0;
			if (np > 0 && acc + np <= SELECT_COMP_SMEM) acc += np;
		}
		cg_off[SELECT_GROUPS] = acc;
	}
	__syncthreads();
	for (int g = 0; g < n_full; ++g) {
		const int o = cg_off[g], np = cg_off[g + 1] - o;
		const size_t slot = ((size_t)blk * SELECT_GROUPS + g) * SELECT_COMP_CAP;
		for (int i = tid; i < np; i += 1024) { cs[o + i] = P.vcomp_start[slot + i]; cd[o + i] = P.vcomp_delta[slot + i]; }
	}
	__syncthreads();
	const uint32_t Q = prefix[nv];
	if (Q > (uint32_t)P.cap) { if (tid == 0) { atomicOr(P.err, 32); P.qcount[blk] = 0; } return; }
	const uint8_t *S1 = P.img + P.blkoff[blk] + 1 + 4 * (size_t)P.m;   // plane-1 snapshot of the block (pbwt.c:298-300)
	const uint32_t m = (uint32_t)P.m;
	// the block's pairs are dealt out over gridDim.y CTAs (each stages the view itself): a launch over few blocks -- one chunk
	// of the load pipeline -- is latency bound, and this is its latency
	for (uint32_t q = blockIdx.y * 1024u + tid; q < Q; q += 1024u * gridDim.y) {
		int lo = 0;                                           // view row of the q-th plane-1 bit: last v with prefix[v] <= q
		for (int len = nv; len > 1;) { const int half = len >> 1; lo += prefix[lo + half] <= q ? half : 0; len -= half; }
		const int v = lo;
		const uint8_t *rec = raw + roff[v];
		uint32_t r = rank_of_one(rec + 9, (uint32_t)(roff[v + 1] - roff[v]) - 9u, q - prefix[v]);
		for (int u = v - 1; u >= 0 && r < m;) {
			const int g = u / COMP_K;
			const int o = cg_off[g], np = cg_off[g + 1] - o;
			if ((u + 1) % COMP_K == 0 && np > 0) {          // a whole group lies behind: one look-up in its inverse composite
				int lo2 = 0;
				for (int len = np; len > 1;) { const int half = len >> 1; lo2 += cs[o + lo2 + half] <= r ? half : 0; len -= half; }
				r += (uint32_t)cd[o + lo2];
				u -= COMP_K;
			} else {
				r = undo_row(raw + roff[u] + 9, (uint32_t)(roff[u + 1] - roff[u]) - 9u, m - n1v[u], r);
				--u;
			}
		}
		const uint32_t col = r < m ? p1_ld_u32_unaligned(S1 + 4 * (size_t)r) : 0xffffffffu;
		if (col >= m) { atomicOr(P.err, 64); continue; }
		P.qcol[(size_t)blk * P.cap + q] = (int32_t)col;
		P.qrow[(size_t)blk * P.cap + q] = P.p1_realrow[vb + v];
	}
	if (tid == 0 && blockIdx.y == 0) P.qcount[blk] = (int)Q;
}

cudaError_t launch_plane1_select(const SelectParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	const size_t smem = sizeof(uint32_t) * (3 * SELECT_MAX_ROWS + 2) + SELECT_MAX_BYTES + 8 * SELECT_COMP_SMEM;
	cudaError_t e = cudaFuncSetAttribute(plane1_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	plane1_select_kernel<<<dim3(n_blk, n_blk <= 32 ? 4 : 1), 1024, smem, st>>>(P);
	return cudaGetLastError();
}

} // namespace b200
