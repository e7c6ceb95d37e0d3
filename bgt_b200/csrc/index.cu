// index.cu -- the row index of a resident PBF, built on the device.
//
// The .pbf only indexes checkpoint blocks (one file offset per 'S' record, pbwt.c:268-276,297); the rows inside a
// block are length-prefixed records ('B', then per plane int32 l + l bytes, pbwt.c:302-308) that pbf_read
// (pbwt.c:313-337) steps through one fread at a time.  Everything per-row that the scan kernels need is derived here
// from the image bytes as they arrive in HBM, so the host only queues the copy:
//   pbf_index_kernel   one warp per block chases the length prefixes through shared-memory rings that are kept full by TMA
//                      bulk copies -> rowoff[blk][0..rows]; validates tags and lengths.  Six lanes share the chain, each
//                      starting at a verified record start of its stretch; lane 0 alone if the pieces do not join up.
//   plan_tiles_kernel  groups rows into the walk kernel's tiles (<= RAW_CAP bytes, never across a COMP_K boundary).
//   p1view_kernel      compacts the rows whose second bit plane is not empty into the "plane-1 view" (a miniature
//                      PBF image per block in fixed-size slots) and decides whether the block qualifies for the split
//                      scan (plane1.cu).
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------ row offsets

constexpr int IX_K = IX_SCRATCH_LANES;         // lanes that chase one block together, each from its own (verified) record start
constexpr int IX_STAGE = 1024, IX_SLOTS = 4, IX_RING = IX_STAGE * IX_SLOTS;   // per-lane ring; small on purpose: a chase CTA
                                                                            // (24 KB) must fit beside the resident CTAs it overlaps
constexpr uint32_t IX_BAD = 0xffffffffu;

__device__ __forceinline__ uint32_t ix_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct IxRing {
	uint8_t  *ring;
	uint64_t *bar;          // [IX_SLOTS]
	const uint8_t *img;
	uint64_t base, end16;   // ring stage k covers image bytes [base + k*IX_STAGE, ...); end16 = end of the copyable range
	long long stage[IX_SLOTS];
	uint32_t parity[IX_SLOTS];
	bool inflight[IX_SLOTS];
	int *err;

	__device__ __forceinline__ void wait(int s)
	{
		uint32_t spins = 0, ok = 0;
		while (true) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(ok) : "r"(ix_smem_u32(bar + s)), "r"(parity[s]) : "memory");
			if (ok) break;
			if (++spins > (1u << 26)) { atomicOr(err, 8); __trap(); }
		}
		parity[s] ^= 1u;
		inflight[s] = false;
	}
	__device__ __forceinline__ void issue(int s, long long k)
	{
		if (inflight[s]) wait(s);
		const uint64_t beg = base + (uint64_t)k * IX_STAGE;
		uint64_t stop = beg + IX_STAGE;
		if (stop > end16) stop = end16;
		stage[s] = k;
		if (beg >= stop) return;
		const uint32_t bytes = (uint32_t)(stop - beg);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ix_smem_u32(bar + s)), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             :: "r"(ix_smem_u32(ring + (size_t)s * IX_STAGE)), "l"(img + beg), "r"(bytes), "r"(ix_smem_u32(bar + s)) : "memory");
		inflight[s] = true;
	}
	__device__ __forceinline__ void want(long long k)
	{
		const int s = (int)(k % IX_SLOTS);
		if (stage[s] != k && base + (uint64_t)k * IX_STAGE < end16) issue(s, k);
	}
	// make stages k and k+1 resident and complete (the chase's window), with the following ones in flight (TMA latency is
	// several stages' worth of chasing); the last slot still holds k-1
	__device__ void window(long long k)
	{
		#pragma unroll 1
		for (int j = 0; j < IX_SLOTS - 1; ++j) want(k + j);
		if (inflight[(int)(k % IX_SLOTS)]) wait((int)(k % IX_SLOTS));
		if (inflight[(int)((k + 1) % IX_SLOTS)]) wait((int)((k + 1) % IX_SLOTS));
	}
};

// the 32-bit little-endian word at ring-relative offset o, any alignment
__device__ __forceinline__ uint32_t ix_read4(uint32_t ring_saddr, uint32_t o)
{
	const uint32_t i0 = o & (uint32_t)(IX_RING - 4), i1 = (i0 + 4) & (uint32_t)(IX_RING - 1);
	uint32_t w0, w1;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(ring_saddr + i0));
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(ring_saddr + i1));
	return __funnelshift_r(w0, w1, (o & 3u) * 8u);
}

// One record at ring-relative offset o: the offset of the next record, or IX_BAD.  Written for latency: offsets are 32-bit
// and relative to the ring base, both length words are read speculatively from the ring (always safe: it is shared
// memory) and ONE combined test decides whether they were resident and valid; only records at the edge of the two-stage
// window take the careful path.
struct IxCursor {
	IxRing R;
	uint32_t ring_saddr, wlo, whi, end;
	uint64_t end64;
	__device__ __forceinline__ uint32_t step(uint32_t o)
	{
		const uint32_t w0 = ix_read4(ring_saddr, o), w1 = ix_read4(ring_saddr, o + 4);
		const uint32_t l0 = __funnelshift_r(w0, w1, 8);
		const uint32_t p1 = o + 5u + l0;
		const uint32_t l1 = ix_read4(ring_saddr, p1);
		const uint32_t nxt = p1 + 4u + l1;
		const bool resident = o >= wlo && o + 8u <= whi && p1 >= wlo && p1 + 8u <= whi;
		const bool valid = (w0 & 0xffu) == 'B' && (l0 | l1) < 0x80000000u && nxt >= p1 && nxt <= end;   // (resident => p1 did not wrap)
		if (resident && valid) return nxt;
		// careful path
		if ((uint64_t)o + 9 > end64) return IX_BAD;
		R.window((long long)(o / IX_STAGE));
		wlo = (o / IX_STAGE) * IX_STAGE; whi = wlo + 2 * IX_STAGE;
		const uint32_t c0 = ix_read4(ring_saddr, o), c1 = ix_read4(ring_saddr, o + 4);
		const uint32_t cl0 = __funnelshift_r(c0, c1, 8);
		const uint64_t cp1 = (uint64_t)o + 5 + cl0;
		if ((c0 & 0xffu) != 'B' || cl0 >= 0x80000000u || cp1 + 4 > end64) return IX_BAD;
		if (!((uint32_t)cp1 >= wlo && (uint32_t)cp1 + 8u <= whi)) {
			R.window((long long)(cp1 / IX_STAGE));
			wlo = (uint32_t)(cp1 / IX_STAGE) * IX_STAGE; whi = wlo + 2 * IX_STAGE;
		}
		const uint32_t cl1 = ix_read4(ring_saddr, (uint32_t)cp1);
		const uint64_t cn = cp1 + 4 + cl1;
		if (cl1 >= 0x80000000u || cn > end64) return IX_BAD;
		return (uint32_t)cn;
	}
};

__device__ __forceinline__ uint32_t ix_gld_u32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

// does a chain of `need` well-formed records start at image offset p (or reach the block end earlier), and do the run
// lengths of both planes of the first one add up to m?  Global loads.  (The length chain alone is not enough: a 'B'
// byte in front of a real length field can fake one record whose end falls on a real boundary -- seen once per ~600
// starts; run lengths of arbitrary bytes do not sum to m.)
__device__ bool ix_verify_start(const uint8_t *img, uint64_t p, uint64_t end, int need, uint32_t m)
{
	{
		if (p + 9 > end || img[p] != 'B') return false;
		const uint32_t l0 = ix_gld_u32(img + p + 1);
		if (l0 >= 0x80000000u || p + 9 + l0 > end) return false;
		const uint32_t l1 = ix_gld_u32(img + p + 5 + l0);
		if (l1 >= 0x80000000u || p + 9 + l0 + (uint64_t)l1 > end) return false;
		unsigned long long t0 = 0, t1 = 0;
		const uint8_t *r0 = img + p + 5, *r1 = img + p + 9 + l0;
		// (a faked record can claim a megabyte of "runs": stop as soon as the sum passes m)
		for (uint32_t i = 0; i < l0 && t0 <= m; ++i) { const uint32_t v = r0[i] >> 1; t0 += (v & 15u) << ((v >> 4) << 2); }
		if (t0 != m) return false;
		for (uint32_t i = 0; i < l1 && t1 <= m; ++i) { const uint32_t v = r1[i] >> 1; t1 += (v & 15u) << ((v >> 4) << 2); }
		if (t1 != m) return false;
	}
	for (int v = 0; v < need; ++v) {
		if (p == end) return v > 0;
		if (p + 9 > end || img[p] != 'B') return false;
		const uint32_t l0 = ix_gld_u32(img + p + 1);
		if (l0 >= 0x80000000u || p + 9 + l0 > end) return false;
		const uint32_t l1 = ix_gld_u32(img + p + 5 + l0);
		if (l1 >= 0x80000000u || p + 9 + l0 + (uint64_t)l1 > end) return false;
		p += 9ull + l0 + l1;
	}
	return true;
}

// One CTA per checkpoint block.  The chain of a block is rows x 2 dependent hops ('B', l0 | plane-0 bytes | l1 | plane-1
// bytes) -- pure latency -- so IX_K chasers share it: chaser 0 starts at the block's first record, chaser l at the first
// position in its stretch of the byte range at which a chain of four well-formed records starts whose first one carries
// run lengths that add up to m in both planes.  A chaser stops where the next one began; if every chaser lands EXACTLY on
// its successor's start and the counts add up to the block's rows, the pieces are the chain and are copied to their
// places.  Otherwise (damaged or unusual blocks) chaser 0 walks the whole block alone.  Every chaser is lane 0 of its own
// warp: in one warp their ring refills (a TMA round trip each kilobyte) would serialise.
__global__ void __launch_bounds__(32 * IX_K) pbf_index_kernel(const IndexParams P)
{
	extern __shared__ __align__(128) uint8_t ix_sm_all[];
	__shared__ uint32_t s_start[IX_K], s_cnt[IX_K];
	__shared__ int s_ok[IX_K];
	const int tid = threadIdx.x, lane = tid >> 5;     // lane = chaser index
	const bool chaser = (tid & 31) == 0;
	const int blk = P.blk_first + (int)blockIdx.x;
	const int BS = 1 << P.shift;
	uint64_t *ro = P.rowoff + (size_t)blk * (BS + 1);
	const int rows = P.rows_in_blk[blk];
	const uint64_t first = P.blkoff[blk] + 1 + 8ull * (uint64_t)P.m;  // behind the 'S' record (pbwt.c:298-300)
	const uint64_t base = first & ~15ull;
	const uint64_t blkend = P.blkend[blk];
	const uint64_t end64 = blkend - base;
	bool bad = P.img[P.blkoff[blk]] != 'S' || first > blkend;
	if (end64 > 0xfff00000ull) { if (tid == 0) atomicOr(P.err, 128); bad = true; }   // the records of one block must fit 32-bit offsets
	if (bad) {
		if (tid == 0) { atomicOr(P.err, 2); P.rows_in_blk[blk] = 0; ro[0] = P.blkoff[blk]; }
		return;
	}
	IxCursor C;
	if (chaser) {
		uint8_t *ix_sm = ix_sm_all + (size_t)lane * (IX_RING + 64);
		C.R.ring = ix_sm; C.R.bar = (uint64_t*)(ix_sm + IX_RING); C.R.img = P.img; C.R.err = P.err;
		C.R.base = base; C.R.end16 = (blkend + 15) & ~15ull;
		#pragma unroll
		for (int s = 0; s < IX_SLOTS; ++s) {
			C.R.stage[s] = -1; C.R.parity[s] = 0; C.R.inflight[s] = false;
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ix_smem_u32(C.R.bar + s)), "r"(1));
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		C.ring_saddr = ix_smem_u32(ix_sm); C.wlo = 1; C.whi = 0; C.end64 = end64; C.end = (uint32_t)end64;
	}
	const uint32_t first_rel = (uint32_t)(first - base), end_rel = (uint32_t)end64;
	const bool team = rows >= 64 * IX_K && P.scratch != nullptr;      // (uniform over the CTA)
	bool done = false;
	if (team) {
		// ---- starts
		if (chaser) {
			uint32_t s = IX_BAD;
			if (lane == 0) s = first_rel;
			else {
				const uint64_t span = blkend - first;
				uint64_t p = first + span * (uint64_t)lane / IX_K;
				const uint64_t lim = first + span * (uint64_t)(lane + 1) / IX_K;
				for (; p < lim; ++p)
					if (P.img[p] == 'B' && ix_verify_start(P.img, p, blkend, 4, (uint32_t)P.m)) { s = (uint32_t)(p - base); break; }
			}
			s_start[lane] = s;
		}
		__syncthreads();
		// ---- chase my stretch into my scratch column; it ends at the first valid start behind me
		if (chaser) {
			const uint32_t s = s_start[lane];
			uint32_t stop = end_rel;
			for (int l = IX_K - 1; l > lane; --l) if (s_start[l] != IX_BAD) stop = s_start[l];
			uint64_t *mine = P.scratch + ((size_t)blk * IX_K + lane) * (size_t)(BS + 1);
			uint32_t cnt = 0, o = s;
			bool ok = true;
			if (s != IX_BAD) {
				while (o < stop) {
					const uint32_t n = C.step(o);
					if (n == IX_BAD || cnt >= (uint32_t)rows) { ok = false; break; }
					mine[cnt++] = base + o;
					o = n;
				}
				if (o != stop) ok = false;              // ran past the successor's start: one of the two is not on the chain
			}
			s_cnt[lane] = cnt; s_ok[lane] = ok ? 1 : 0;
		}
		__syncthreads();                                // (also makes the scratch columns visible to the whole CTA)
		bool all_ok = true;
		uint32_t total = 0;
		for (int l = 0; l < IX_K; ++l) { all_ok = all_ok && s_ok[l]; total += s_cnt[l]; }
		if (all_ok && total == (uint32_t)rows) {
			// ---- the pieces are the chain: copy them to their places (coalesced, all threads)
			uint32_t b0 = 0;
			for (int l = 0; l < IX_K; ++l) {
				const uint32_t c = s_cnt[l];
				const uint64_t *src = P.scratch + ((size_t)blk * IX_K + l) * (size_t)(BS + 1);
				for (uint32_t i = tid; i < c; i += 32 * IX_K) ro[b0 + i] = src[i];
				b0 += c;
			}
			if (tid == 0) ro[rows] = blkend;
			done = true;
		}
	}
	if (!done && tid == 0) { // ---- the whole block, alone
		if (team) atomicAdd(P.fallbacks, 1);
		uint32_t o = first_rel;
		int r = 0;
		for (; r < rows; ++r) {
			const uint32_t n = C.step(o);
			if (n == IX_BAD) { bad = true; break; }
			ro[r] = base + o;
			o = n;
		}
		if (!bad) ro[rows] = base + o;
	}
	if (chaser) {
		#pragma unroll 1
		for (int s = 0; s < IX_SLOTS; ++s) if (C.R.inflight[s]) C.R.wait(s);   // no copy may be in flight when the CTA exits
	}
	if (bad && tid == 0) {
		// the records of this block do not parse: it decodes as an empty block and the load reports the corruption
		atomicOr(P.err, 2);
		P.rows_in_blk[blk] = 0;
		ro[0] = P.blkoff[blk];
	}
}

cudaError_t launch_index(const IndexParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	const size_t smem = (size_t)IX_K * (IX_RING + 64);
	cudaError_t e = cudaFuncSetAttribute(pbf_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	pbf_index_kernel<<<n_blk, 32 * IX_K, smem, st>>>(P);
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ tiles

// Tiles: maximal groups of consecutive rows with <= T_MAX rows and <= RAW_CAP bytes that do not cross a multiple of
// COMP_K rows; a row larger than RAW_CAP on its own is a single-row "big" tile (streamed in pieces by the walk).
// Tiles of different COMP_K-row groups are independent: one thread per group, two passes (count, place).
__device__ __forceinline__ int tiles_of_group(const uint64_t *ro, int r, int r_end, int2 *out)
{
	int n = 0;
	while (r < r_end) {
		if (ro[r + 1] - ro[r] > (uint64_t)RAW_CAP) { if (out) out[n] = make_int2(r, (int)(1u | 0x80000000u)); ++n; ++r; continue; }
		int e = r + 1;
		while (e < r_end && e - r < T_MAX && ro[e + 1] - ro[r] <= (uint64_t)RAW_CAP) ++e;
		if (out) out[n] = make_int2(r, e - r);
		++n;
		r = e;
	}
	return n;
}

__global__ void __launch_bounds__(256) plan_tiles_kernel(const IndexParams P)
{
	__shared__ int warp_tot[8];
	__shared__ int s_base;
	const int blk = P.blk_first + (int)blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int BS = 1 << P.shift;
	const int n_grp = (BS + COMP_K - 1) / COMP_K;
	const uint64_t *ro = P.rowoff + (size_t)blk * (BS + 1);
	const int rows = P.rows_in_blk[blk];
	const int t0 = blk * BS;                       // the block's tiles live in a fixed slot of BS entries
	int *gtb = P.grp_tile_beg + (size_t)blk * (n_grp + 1);
	if (tid == 0) s_base = 0;
	__syncthreads();
	for (int g0 = 0; g0 < n_grp; g0 += 256) {
		const int g = g0 + tid;
		int r_lo = g * COMP_K, r_hi = r_lo + COMP_K;
		if (r_hi > rows) r_hi = rows;
		const int mine = (g < n_grp && r_lo < rows) ? tiles_of_group(ro, r_lo, r_hi, nullptr) : 0;
		int x = mine;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
		if (lane == 31) warp_tot[warp] = x;
		__syncthreads();
		int before = s_base + x - mine;
		for (int w = 0; w < warp; ++w) before += warp_tot[w];
		if (g < n_grp) {
			gtb[g] = t0 + before;
			if (mine) tiles_of_group(ro, r_lo, r_hi, P.tiles + t0 + before);
		}
		__syncthreads();
		if (tid == 255) s_base = before + mine;
		__syncthreads();
	}
	if (tid == 0) { gtb[n_grp] = t0 + s_base; P.blk_tile_beg[blk] = t0; P.blk_tile_end[blk] = t0 + s_base; }
}

cudaError_t launch_plan_tiles(const IndexParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	plan_tiles_kernel<<<n_blk, 256, 0, st>>>(P);
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ plane-1 view

__device__ __forceinline__ uint32_t ix_ld_u32_unaligned(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	const uint32_t lo = w[0];
	if (sh == 0) return lo;
	return __funnelshift_r(lo, w[1], sh);
}

constexpr int P1V_NT = 256;

// Per block: the rows whose plane 1 has at least one 1 bit (n1 from rowmeta_kernel; a corrupt row counts as empty),
// re-framed as records 'B', l0 = 0, l1, bytes -- an empty plane 0 -- in the block's slot of the view image, with their
// n1 and row number.  A block is "sparse" (split scan applies) if its view fits the select kernel's staging buffers.
__global__ void __launch_bounds__(P1V_NT) p1view_kernel(const P1ViewParams P)
{
	__shared__ uint32_t wrow[P1V_NT / 32], wbyte[P1V_NT / 32];
	__shared__ uint32_t s_rows, s_bytes, s_allones;
	__shared__ unsigned long long s_ones;
	const int blk = P.blk_first + (int)blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int BS = 1 << P.shift;
	const uint64_t *ro = P.rowoff + (size_t)blk * (BS + 1);
	const int rows = P.rows_in_blk[blk];
	const uint32_t m = (uint32_t)P.m;
	const size_t vb = (size_t)blk * SELECT_MAX_ROWS;
	const uint64_t slot = (uint64_t)blk * P1_SLOT_BYTES;
	uint64_t *vro = P.p1_rowoff + vb + blk;
	uint32_t *vpre = P.p1_prefix + vb + blk;
	__shared__ uint32_t wones[P1V_NT / 32];
	__shared__ uint32_t s_pre;
	if (tid == 0) { s_rows = 0; s_bytes = 0; s_allones = 0; s_ones = 0; s_pre = 0; }
	__syncthreads();
	for (int r0 = 0; r0 < rows; r0 += P1V_NT) {
		const int r = r0 + tid;
		uint32_t n1 = 0, l1 = 0;
		const uint8_t *rle = nullptr;
		if (r < rows) {
			n1 = P.n1[((size_t)blk * BS + r) * 2 + 1];
			if (n1) {
				const uint8_t *rec = P.img + ro[r];
				const uint32_t l0 = ix_ld_u32_unaligned(rec + 1);
				l1 = ix_ld_u32_unaligned(rec + 5 + l0);
				rle = rec + 9 + l0;
			}
		}
		const uint32_t f = n1 ? 1u : 0u, sz = n1 ? 9u + l1 : 0u;
		const uint32_t n1c = n1 < (1u << 24) ? n1 : (1u << 24);   // (a block whose plane 1 holds that many ones is not sparse anyway; keeps the prefix from wrapping)
		uint32_t x = f, y = sz, z = n1c;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t tx = __shfl_up_sync(0xffffffffu, x, d), ty = __shfl_up_sync(0xffffffffu, y, d), tz = __shfl_up_sync(0xffffffffu, z, d);
			if (lane >= d) { x += tx; y += ty; z += tz; }
		}
		if (lane == 31) { wrow[warp] = x; wbyte[warp] = y; wones[warp] = z; }
		__syncthreads();
		uint32_t v = s_rows + x - f, off = s_bytes + y - sz, pre = s_pre + z - n1c;
		for (int w = 0; w < warp; ++w) { v += wrow[w]; off += wbyte[w]; pre += wones[w]; }
		if (n1) {
			if (n1 == m) s_allones = 1;
			atomicAdd(&s_ones, (unsigned long long)n1);
			if (v < (uint32_t)SELECT_MAX_ROWS && off + sz <= (uint32_t)SELECT_MAX_BYTES) {
				vro[v] = slot + off;
				vpre[v] = pre;
				P.p1_n1[vb + v] = n1;
				P.p1_realrow[vb + v] = (uint16_t)r;
				uint8_t *dst = P.p1img + slot + off;
				dst[0] = 'B'; dst[1] = dst[2] = dst[3] = dst[4] = 0;
				dst[5] = (uint8_t)l1; dst[6] = (uint8_t)(l1 >> 8); dst[7] = (uint8_t)(l1 >> 16); dst[8] = (uint8_t)(l1 >> 24);
				for (uint32_t i = 0; i < l1; ++i) dst[9 + i] = rle[i];
			}
		}
		__syncthreads();
		if (tid == P1V_NT - 1) { s_rows = v + f; s_bytes = off + sz; s_pre = pre + n1c; }
		__syncthreads();
	}
	if (tid == 0) {
		const bool fits = s_rows <= (uint32_t)SELECT_MAX_ROWS && s_bytes <= (uint32_t)SELECT_MAX_BYTES;
		const bool sparse = fits && !s_allones && s_ones <= (unsigned long long)P.p1_cap && BS <= 65536;
		P.blk_sparse[blk] = sparse ? (s_ones <= (unsigned long long)P.p1_base ? 1 : 2) : 0;
		if (P.blk_ones) P.blk_ones[blk] = s_ones < 0xffffffffull ? (uint32_t)s_ones : 0xffffffffu;
		P.p1_rows_in_blk[blk] = fits ? (int)s_rows : 0;
		P.p1_vbase[blk] = (long long)vb;
		vro[fits ? s_rows : 0] = slot + (fits ? s_bytes : 0);
		vpre[fits ? s_rows : 0] = fits ? s_pre : 0u;
		if (fits) { // zero padding behind the records (the view is read in aligned words)
			uint8_t *dst = P.p1img + slot + s_bytes;
			for (int i = 0; i < 16; ++i) dst[i] = 0;
		} else vro[0] = slot;
	}
}

cudaError_t launch_p1view(const P1ViewParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	p1view_kernel<<<n_blk, P1V_NT, 0, st>>>(P);
	return cudaGetLastError();
}

} // namespace b200
