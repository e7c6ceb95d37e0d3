// encode.cu -- the PBWT encoder on the device (pbf_write, pbwt.c:288-311; pbc_enc_core, pbwt.c:57-66; pbr_enc,
// pbwt.c:24-50).
//
// The reference gathers the row through the whole permutation (u[j] = a[S0[j]]), partitions S0 into S and run-length
// encodes u -- 4m bytes of permutation read and written per row and plane.  Here it is the rank walk run forward, like
// decode: every column keeps its RANK per plane in a register; a row is encoded by
//   scatter   every column ORs its bit into a bit vector at its rank                        (global atomics, L2)
//   barrier   one grid-wide barrier per row (cooperative launch; the kernel is persistent over a batch of rows)
//   rank      every CTA stages the m-bit vector in shared memory with a popcount prefix per 4 words; a column's new rank
//             is (#0s before it) for a 0 bit and (#0s total + #1s before it) for a 1 bit       (pbwt.c:62-64)
//   RLE       one CTA per plane turns the bit vector into the run-length byte code: run starts from word XORs, the
//             distance to the next start by a suffix-min over words, byte counts by a prefix sum, then the hex-digit
//             bytes of pbr_enc1 (most significant digit first, one byte per non-zero digit)    (pbwt.c:24-36)
// and every 2^shift rows all columns dump S[rank] = column (the 'S' snapshot, pbwt.c:292-301).
// Rows of a batch arrive as bit planes in column order; the byte-per-haplotype rows of pbf_write are packed on the way in.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace cg = cooperative_groups;

namespace b200 {

constexpr int ENC_NT = 1024, ENC_NW = ENC_NT / 32, ENC_SEG = 4096;

__device__ __forceinline__ uint32_t enc_nz_nibbles(uint32_t L) // number of non-zero hex digits = bytes of pbr_enc1
{
	const uint32_t x = (L | L >> 1 | L >> 2 | L >> 3) & 0x11111111u;
	return (uint32_t)__popc(x);
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = sum (all threads)
__device__ __forceinline__ uint32_t enc_block_scan(uint32_t v, uint32_t *warp_tot, uint32_t *total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t x = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
	__syncthreads();                       // warp_tot may still be read from the previous use
	if (lane == 31) warp_tot[warp] = x;
	__syncthreads();
	uint32_t before = x - v, tot = 0;
	for (int w = 0; w < ENC_NW; ++w) { const uint32_t t = warp_tot[w]; if (w < warp) before += t; tot += t; }
	*total = tot;
	return before;
}

template<int C>
__global__ void __launch_bounds__(ENC_NT, 1) pbwt_encode_kernel(const EncodeParams P)
{
	extern __shared__ __align__(16) uint8_t enc_sm[];
	cg::grid_group grid = cg::this_grid();
	const int words = P.words, w4 = (words + 3) / 4;
	uint32_t *V = (uint32_t*)enc_sm;                 // [words + 4] the row's bit vector in rank order
	uint32_t *pre4 = V + ((words + 4 + 3) & ~3);     // [w4 + 1] ones before every group of 4 words
	uint32_t *nxt = pre4 + ((w4 + 1 + 3) & ~3);      // [ENC_SEG + 1] emitter: first run start at or behind a word
	uint32_t *boff = nxt + ENC_SEG + 4;              // [ENC_SEG + 1] emitter: byte offset of a word's runs
	__shared__ uint32_t warp_tot[ENC_NW];
	__shared__ uint32_t seg_first[64];               // emitter: first run start of every segment (m <= 64 * 32 * ENC_SEG)
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t m = (uint32_t)P.m;
	const int n_cta = (int)gridDim.x;
	const int slice = (int)blockIdx.x * (ENC_NT * C);
	const int BSm1 = (1 << P.shift) - 1;

	uint32_t rk[2][C];
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t col = (uint32_t)(slice + c * ENC_NT + tid);
		rk[0][c] = col < m ? (uint32_t)P.rank[col] : 0u;
		rk[1][c] = col < m ? (uint32_t)P.rank[(size_t)m + col] : 0u;
	}
	unsigned long long out_pos[2] = {P.out_pos0[0], P.out_pos0[1]};   // emitters: running offsets into the plane streams

	for (int r = 0; r < P.n_rows; ++r) {
		const long long arow = P.row0 + r;
		const int buf = (int)(arow % 3);
		uint32_t *B = P.bitvec + (size_t)buf * 2 * words;
		// ---- snapshot (pbwt.c:292-301): S[rank] = column, both planes
		if ((arow & BSm1) == 0) {
			int32_t *S = P.snap + (size_t)((arow >> P.shift) - (P.row0 + BSm1 >> P.shift)) * 2 * (size_t)m;
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const uint32_t col = (uint32_t)(slice + c * ENC_NT + tid);
				if (col < m) { S[rk[0][c]] = (int32_t)col; S[(size_t)m + rk[1][c]] = (int32_t)col; }
			}
		}
		// ---- scatter the row's bits to rank order; clear the vector of the next row
		{
			const uint32_t *in = P.in_bits + (size_t)r * 2 * words;
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const uint32_t col = (uint32_t)(slice + c * ENC_NT + tid);
				if (col < m) {
					if ((in[col >> 5] >> (col & 31)) & 1u) atomicOr(&B[rk[0][c] >> 5], 1u << (rk[0][c] & 31));
					if ((in[words + (col >> 5)] >> (col & 31)) & 1u) atomicOr(&B[words + (rk[1][c] >> 5)], 1u << (rk[1][c] & 31));
				}
			}
			uint32_t *Bn = P.bitvec + (size_t)((buf + 1) % 3) * 2 * words;
			for (int i = (int)blockIdx.x * ENC_NT + tid; i < 2 * words; i += n_cta * ENC_NT) Bn[i] = 0u;
		}
		grid.sync();
		#pragma unroll
		for (int p = 0; p < 2; ++p) {
			// ---- stage the vector; ones before every 4 words
			const uint32_t *Bp = B + (size_t)p * words;
			for (int w = tid; w < w4 * 4; w += ENC_NT) V[w] = w < words ? __ldcg(Bp + w) : 0u;
			__syncthreads();
			// ones before every group of 4 words: one block scan per pass of ENC_NT groups
			uint32_t n1;
			{
				uint32_t base = 0;
				for (int g0 = 0; g0 < w4; g0 += ENC_NT) {
					const int g4 = g0 + tid;
					uint32_t s = 0;
					if (g4 < w4) {
						#pragma unroll
						for (int j = 0; j < 4; ++j) s += (uint32_t)__popc(V[g4 * 4 + j]);
					}
					uint32_t tot;
					const uint32_t before = enc_block_scan(s, warp_tot, &tot);
					if (g4 < w4) pre4[g4] = base + before;
					base += tot;
				}
				n1 = base;
			}
			__syncthreads();
			const uint32_t zeros = m - n1;
			// ---- new ranks (pbwt.c:62-64: 0s keep their order in front, 1s behind)
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const uint32_t col = (uint32_t)(slice + c * ENC_NT + tid);
				if (col < m) {
					const uint32_t q = rk[p][c], w = q >> 5;
					uint32_t ones = pre4[w >> 2];
					for (uint32_t j = w & ~3u; j < w; ++j) ones += (uint32_t)__popc(V[j]);
					ones += (uint32_t)__popc(V[w] & ((1u << (q & 31)) - 1u));
					rk[p][c] = ((V[w] >> (q & 31)) & 1u) ? zeros + ones : q - ones;
				}
			}
			// ---- run-length code of the vector, by the plane's emitter CTA (pbwt.c:24-50)
			if ((int)blockIdx.x == p % n_cta) {
				uint8_t *out = P.out[p] + out_pos[p];
				const int n_seg = (words + ENC_SEG - 1) / ENC_SEG;
				// run starts of word w: bit i set = a run starts at rank 32w+i
				auto starts_of = [&](int w) -> uint32_t {
					const uint32_t v = V[w];
					const uint32_t prev = w ? V[w - 1] >> 31 : (~v & 1u);           // rank 0 always starts a run
					uint32_t t = v ^ (v << 1 | prev);
					const uint32_t hi = m - (uint32_t)w * 32u;                       // valid bits in this word
					if (hi < 32u) t &= (1u << hi) - 1u;
					return t;
				};
				// first start of every segment
				if (tid < 64) seg_first[tid] = m;
				__syncthreads();
				for (int w = tid; w < words; w += ENC_NT) {
					const uint32_t t = starts_of(w);
					if (t) atomicMin(&seg_first[w / ENC_SEG], (uint32_t)w * 32u + (uint32_t)__ffs(t) - 1u);
				}
				__syncthreads();
				uint32_t l_total = 0;
				for (int sg = 0; sg < n_seg; ++sg) {
					const int w_lo = sg * ENC_SEG, w_hi = w_lo + ENC_SEG < words ? w_lo + ENC_SEG : words, nw = w_hi - w_lo;
					uint32_t after = m;                                              // first start behind this segment
					for (int s2 = n_seg - 1; s2 > sg; --s2) if (seg_first[s2] < m) after = seg_first[s2];
					// nxt[i] = first start at or behind word w_lo+i (suffix minimum); contiguous chunk per thread
					const int per = (nw + ENC_NT - 1) / ENC_NT, c_lo = tid * per, c_hi = c_lo + per < nw ? c_lo + per : nw;
					uint32_t mine = m;
					for (int i = c_hi - 1; i >= c_lo; --i) {
						const uint32_t t = starts_of(w_lo + i);
						if (t) mine = (uint32_t)(w_lo + i) * 32u + (uint32_t)__ffs(t) - 1u;
						nxt[i] = mine;
					}
					// first start behind my chunk: suffix minimum over the later chunks (starts grow with the chunk index), else `after`
					uint32_t x = mine;
					#pragma unroll
					for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_down_sync(0xffffffffu, x, d); if (lane + d < 32 && t < x) x = t; }
					__syncthreads();
					if (lane == 0) warp_tot[tid >> 5] = x;                          // first start in my warp's chunks
					__syncthreads();
					uint32_t behind = after;
					for (int w = ENC_NW - 1; w > (tid >> 5); --w) if (warp_tot[w] < m) behind = warp_tot[w];
					{
						const uint32_t nb = __shfl_down_sync(0xffffffffu, x, 1);    // first start in the chunks of the lanes behind me
						if (lane < 31 && nb < m) behind = nb;
					}
					for (int i = c_hi - 1; i >= c_lo && nxt[i] == m; --i) nxt[i] = behind;
					if (tid == 0) nxt[nw] = after;
					__syncthreads();
					// bytes per word
					uint32_t base = 0;
					for (int i0 = 0; i0 < nw; i0 += ENC_NT) {
						const int i = i0 + tid;
						uint32_t nb = 0;
						if (i < nw) {
							uint32_t t = starts_of(w_lo + i);
							while (t) {
								const uint32_t b = (uint32_t)__ffs(t) - 1u;
								t &= t - 1u;
								const uint32_t pos = (uint32_t)(w_lo + i) * 32u + b;
								const uint32_t next = t ? (uint32_t)(w_lo + i) * 32u + (uint32_t)__ffs(t) - 1u : nxt[i + 1];
								nb += enc_nz_nibbles(next - pos);
							}
						}
						uint32_t tot;
						const uint32_t before = enc_block_scan(nb, warp_tot, &tot);
						if (i < nw) boff[i] = base + before;
						base += tot;
					}
					__syncthreads();
					// the bytes
					for (int i = tid; i < nw; i += ENC_NT) {
						uint32_t t = starts_of(w_lo + i);
						uint8_t *q = out + l_total + boff[i];
						while (t) {
							const uint32_t b = (uint32_t)__ffs(t) - 1u;
							t &= t - 1u;
							const uint32_t pos = (uint32_t)(w_lo + i) * 32u + b;
							const uint32_t next = t ? (uint32_t)(w_lo + i) * 32u + (uint32_t)__ffs(t) - 1u : nxt[i + 1];
							const uint32_t L = next - pos, bit = (V[w_lo + i] >> b) & 1u;
							for (int d = 7; d >= 0; --d) {
								const uint32_t dig = (L >> (4 * d)) & 15u;
								if (dig) *q++ = (uint8_t)((((uint32_t)d << 4 | dig) << 1) | bit);
							}
						}
					}
					l_total += base;
					__syncthreads();
				}
				if (tid == 0) P.row_len[(size_t)r * 2 + p] = l_total;
				out_pos[p] += l_total;
			}
			__syncthreads();
		}
	}
	// ---- state back to HBM for the next batch
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t col = (uint32_t)(slice + c * ENC_NT + tid);
		if (col < m) { P.rank[col] = (int32_t)rk[0][c]; P.rank[(size_t)m + col] = (int32_t)rk[1][c]; }
	}
	#pragma unroll
	for (int p = 0; p < 2; ++p)
		if ((int)blockIdx.x == p % n_cta && tid == 0) P.out_pos1[p] = out_pos[p];
}

size_t encode_smem_bytes(int m)
{
	const int words = (m + 31) / 32, w4 = (words + 3) / 4;
	return sizeof(uint32_t) * (size_t)(((words + 4 + 3) & ~3) + ((w4 + 1 + 3) & ~3) + 2 * (ENC_SEG + 4));
}

template<int C>
static cudaError_t launch_encode_t(const EncodeParams &P, int n_cta, size_t smem, cudaStream_t st)
{
	cudaError_t e = cudaFuncSetAttribute(pbwt_encode_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	void *args[] = {(void*)&P};
	return cudaLaunchCooperativeKernel((void*)pbwt_encode_kernel<C>, dim3(n_cta), dim3(ENC_NT), args, smem, st);
}

// all CTAs must be co-resident (grid barrier): columns per thread grow until the grid fits one CTA per SM
cudaError_t launch_encode(const EncodeParams &P, int sm_count, cudaStream_t st)
{
	if (P.n_rows <= 0) return cudaSuccess;
	const size_t smem = encode_smem_bytes(P.m);
	if (smem > 220 * 1024 || P.words > 64 * ENC_SEG) return cudaErrorInvalidValue;
	int C = 1;
	while (C < 8 && (P.m + ENC_NT * C - 1) / (ENC_NT * C) > sm_count) C <<= 1;
	const int n_cta = (P.m + ENC_NT * C - 1) / (ENC_NT * C);
	if (n_cta > sm_count) return cudaErrorInvalidValue;
	switch (C) {
	case 1: return launch_encode_t<1>(P, n_cta, smem, st);
	case 2: return launch_encode_t<2>(P, n_cta, smem, st);
	case 4: return launch_encode_t<4>(P, n_cta, smem, st);
	default: return launch_encode_t<8>(P, n_cta, smem, st);
	}
}

// pbf_write's rows (one byte per haplotype, pbwt.c:61 takes !!a[]) -> bit planes in column order
__global__ void __launch_bounds__(256) pack_rows_kernel(const uint8_t *__restrict__ a0, const uint8_t *__restrict__ a1, long long n_rows, int m, int words,
                                                        uint32_t *__restrict__ bits)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rows * 2 * words) return;
	const int w = (int)(i % words), p = (int)((i / words) & 1);
	const long long r = i / (2 * words);
	const uint8_t *src = (p ? a1 : a0) + (size_t)r * m + (size_t)w * 32;
	const int n = m - w * 32 < 32 ? m - w * 32 : 32;
	uint32_t v = 0;
	for (int k = 0; k < n; ++k) v |= (src[k] ? 1u : 0u) << k;
	bits[i] = v;
}

cudaError_t launch_pack_rows(const uint8_t *a0, const uint8_t *a1, long long n_rows, int m, uint32_t *bits, cudaStream_t st)
{
	const int words = (m + 31) / 32;
	const long long n = n_rows * 2 * words;
	if (n <= 0) return cudaSuccess;
	pack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a0, a1, n_rows, m, words, bits);
	return cudaGetLastError();
}

} // namespace b200
