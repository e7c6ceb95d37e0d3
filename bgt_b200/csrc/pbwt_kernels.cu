// pbwt_kernels.cu -- hand-written sm_100a kernels of BGT's genotype hot path.
//
// What the reference does per site (row) on one CPU thread:
//   pbf_read (pbwt.c:313-337) -> pbc_dec_core (pbwt.c:69-90) per bit plane: rebuild the whole permutation S_k
//   from S_{k-1} and scatter the row's bits to column order; then bgtm_cal_info (bgt.c:735-757) walks the
//   2*n_out decoded bytes and histograms the 2-bit code per sample group.
// What this file does instead: the permutation is never materialised.  Every tracked column
// keeps its own RANK in the current PBWT order of each plane (rank = S^-1[column]); a row only needs the run
// that contains that rank:   bit = run's bit,   rank' = rank + delta(run)           (pbwt.c:142-153, the
// arithmetic of pbs_dec applied to every column).  Columns never interact, rows of one checkpoint block are
// walked in order by the thread that owns the column, and checkpoint blocks (pbwt.c:292-301) are independent.
//
//   grid  = (column slices, checkpoint blocks);  CTA = 512 threads, C columns per thread, both planes.
//   tile  = up to 64 consecutive rows whose 'B' records fit 8 KB: ONE TMA bulk copy (cp.async.bulk +
//           mbarrier) brings the raw RLE bytes of the next tile into shared memory while the current tile is
//           being walked.
//   parse = one warp per (row, plane): byte -> (length, bit) by arithmetic (the 128-entry table of
//           pbwt.c:12-21 is (v&15)<<4*(v>>4)), two warp-shuffle scans give every byte's start rank and the
//           rank shift delta of its run; tables stay in shared memory.
//   walk  = per column a branch-free binary search over the run starts (shared memory), rank += delta,
//           bit = (rank' >= #zeros).
//   count = per 32 columns one __ballot_sync per plane; lane g < G holds group g's membership mask and popcounts
//           ALT / missing / other-ALT (bgt.c:743-756) -- per-row shared-memory counters, flushed once per tile.
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

#define FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------------ helpers

__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p)
{
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t*)(a & ~(uintptr_t)3);
	const uint32_t sh = (uint32_t)(a & 3) * 8;
	const uint32_t lo = w[0];
	if (sh == 0) return lo;
	return __funnelshift_r(lo, w[1], sh);
}

__device__ __forceinline__ uint32_t lds_u32_bytes(const uint8_t *p) // shared memory, any alignment
{
	return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

// pbwt.c:12-21 as arithmetic: byte c -> run length
__device__ __forceinline__ uint32_t rle_len(uint32_t c)
{
	const uint32_t v = c >> 1;
	return (v & 15u) << ((v >> 4) << 2);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}

// 1-D TMA bulk copy global -> shared of this CTA; completion is signalled on the mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ row meta

// One warp per row: n1 (number of 1 bits) of both planes, a check that the run lengths sum to m, and the number of runs
// of plane 0 (bytes of the same symbol merged, empty bytes skipped) -- what compose.cu needs to place the rows' maps.
__global__ void __launch_bounds__(256) rowmeta_kernel(const uint8_t *__restrict__ img, const uint64_t *__restrict__ rowoff,
                                                      int n_blk, int shift, const int *__restrict__ rows_in_blk, uint32_t m,
                                                      uint32_t *__restrict__ n1, uint32_t *__restrict__ nrun0, unsigned long long *__restrict__ bad)
{
	const int lane = threadIdx.x & 31;
	const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int BS = 1 << shift;
	const int blk = (int)(wid >> shift), r = (int)(wid & (BS - 1));
	if (blk >= n_blk || r >= rows_in_blk[blk]) return;
	const uint8_t *p = img + rowoff[(size_t)blk * (BS + 1) + r] + 1; // skip 'B'
	for (int plane = 0; plane < 2; ++plane) {
		const uint32_t l = ld_u32_unaligned(p);
		p += 4;
		unsigned long long tot = 0, ones = 0;
		uint32_t nrun = 0, prev_bit = 2, first_bit = 0;
		for (uint32_t base = 0; base < l; base += 32) {
			const uint32_t i = base + lane;
			const uint32_t c = i < l ? p[i] : 0u, len = rle_len(c), b = c & 1u;
			tot += len;
			if (b) ones += len;
			if (plane == 0 && nrun0) {
				const uint32_t valid = __ballot_sync(FULL_MASK, len > 0), bitm = __ballot_sync(FULL_MASK, b != 0);
				const uint32_t below = valid & ((1u << lane) - 1u);
				const uint32_t pb = below ? (bitm >> (31 - __clz(below))) & 1u : prev_bit;
				if (prev_bit == 2 && valid) first_bit = (bitm >> (__ffs(valid) - 1)) & 1u;
				nrun += __popc(__ballot_sync(FULL_MASK, len > 0 && pb != b));
				if (valid) prev_bit = (bitm >> (31 - __clz(valid))) & 1u;
			}
		}
		#pragma unroll
		for (int d = 16; d; d >>= 1) {
			tot += __shfl_xor_sync(FULL_MASK, tot, d);
			ones += __shfl_xor_sync(FULL_MASK, ones, d);
		}
		if (lane == 0) {
			const bool ok = (tot == m);
			n1[((size_t)blk * BS + r) * 2 + plane] = ok ? (uint32_t)ones : 0u; // a corrupt row decodes as all-REF
			if (!ok) atomicAdd(bad, 1ull);
			if (plane == 0 && nrun0) nrun0[(size_t)blk * BS + r] = (nrun < (1u << 20) ? nrun : (1u << 20)) | first_bit << 31;   // compose.cu places the rows' maps with this
		}
		p += l;
	}
}

cudaError_t launch_rowmeta(const uint8_t *img, const uint64_t *rowoff, int n_blk, int shift, long long, const int *rows_in_blk,
                           uint32_t m, uint32_t *n1, uint32_t *nrun0, unsigned long long *bad, cudaStream_t st)
{
	const long long warps = (long long)n_blk << shift;
	if (warps == 0) return cudaSuccess;
	const long long blocks = (warps * 32 + 255) / 256;
	rowmeta_kernel<<<(unsigned)blocks, 256, 0, st>>>(img, rowoff, n_blk, shift, rows_in_blk, m, n1, nrun0, bad);
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ snapshots -> ranks

// pbwt.c:343 (invS[S[i]] = i) for every resident block and plane.  The int32 arrays sit at odd file offsets.
__global__ void __launch_bounds__(256) invert_snapshot_kernel(const uint8_t *__restrict__ img, const uint64_t *__restrict__ blkoff,
                                                              int m, int32_t *__restrict__ rank0, int *__restrict__ err)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const int blk = blockIdx.y, plane = blockIdx.z;
	const uint8_t *S = img + blkoff[blk] + 1 + (size_t)plane * 4 * (size_t)m;
	const uint32_t c = ld_u32_unaligned(S + 4 * (size_t)i);
	if (c < (uint32_t)m) rank0[((size_t)blk * 2 + plane) * (size_t)m + c] = i;
	else atomicOr(err, 1);
}

cudaError_t launch_invert_snapshots(const uint8_t *img, const uint64_t *blkoff, int n_blk, int m, int32_t *rank0, int *err, cudaStream_t st)
{
	if (n_blk == 0 || m == 0) return cudaSuccess;
	for (int b0 = 0; b0 < n_blk; b0 += 32768) { // gridDim.y limit 65535
		const int nb = n_blk - b0 < 32768 ? n_blk - b0 : 32768;
		dim3 grid((m + 255) / 256, nb, 2);
		invert_snapshot_kernel<<<grid, 256, 0, st>>>(img, blkoff + b0, m, rank0 + (size_t)b0 * 2 * (size_t)m, err);
	}
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ the rank walk

// One warp turns `len` RLE bytes (shared memory, at raw+off) into the run table of the row: for byte i
//   ts[off+i] = rank at which its run starts,  td[off+i] = new_rank - rank for ranks inside the run.
// (tot0, ones0) = rank / number of ones already consumed (non-zero only when a long row is parsed in pieces).
// Zero-length bytes (legal on input) get the start of the following run and are never hit.
__device__ __forceinline__ void parse_runs(const uint8_t *raw, uint32_t off, uint32_t len, uint32_t zeros_total,
                                           uint32_t *ts, int32_t *td, int lane, uint32_t &tot0, uint32_t &ones0)
{
	uint32_t tot = tot0, ones = ones0;
	for (uint32_t base = 0; base < len; base += 32) {
		const uint32_t i = base + lane;
		const uint32_t c = i < len ? raw[off + i] : 0u;
		const uint32_t L = rle_len(c), b = c & 1u, L1 = b ? L : 0u;
		uint32_t x = L, y = L1;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t tx = __shfl_up_sync(FULL_MASK, x, d), ty = __shfl_up_sync(FULL_MASK, y, d);
			if (lane >= d) { x += tx; y += ty; }
		}
		const uint32_t start = tot + x - L;               // ranks before this run
		const uint32_t ones_before = ones + y - L1;       // ...of which in 1-runs
		const uint32_t zeros_before = start - ones_before;
		// pbwt.c:150: rank' = acc[b] + c[b] + (rank - s);  acc = {0, m-n1}
		const int32_t delta = b ? (int32_t)(zeros_total - zeros_before) : -(int32_t)ones_before;
		if (i < len) { ts[off + i] = start; td[off + i] = delta; }
		tot += __shfl_sync(FULL_MASK, x, 31);
		ones += __shfl_sync(FULL_MASK, y, 31);
	}
	tot0 = tot; ones0 = ones;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
	return v;
}

// byte distance between a run's start (ts[]) and its delta (td[]) in shared memory: the two tables are adjacent
constexpr uint32_t TD_MINUS_TS = sizeof(uint32_t) * RAW_BYTES;

// C independent branch-free binary searches over the n run starts at shared address ts_saddr: last run whose
// start <= rank.  The cursor is kept as a shared-memory ADDRESS so one level costs add / LDS / compare / select.
template<int C>
__device__ __forceinline__ void lookup_runs(uint32_t (&r)[C], uint32_t ts_saddr, uint32_t n, uint32_t zeros_total, uint32_t &bits)
{
	// (the window halves as len - len/2, not by powers of two: power-of-two strides put the probes of all lanes into the
	// same shared-memory bank from the sixth level on -- measured 1.45x slower)
	uint32_t a[C];
	#pragma unroll
	for (int c = 0; c < C; ++c) a[c] = ts_saddr;
	for (uint32_t len = n; len > 1;) {
		const uint32_t half = len >> 1, h4 = half << 2;
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const uint32_t t = a[c] + h4;
			a[c] = lds_u32(t) <= r[c] ? t : a[c];
		}
		len -= half;
	}
	uint32_t b = 0;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		r[c] += lds_u32(a[c] + TD_MINUS_TS);
		b |= (r[c] >= zeros_total ? 1u : 0u) << c;
	}
	bits = b;
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr)
{
	uint16_t v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
	return v;
}

// Look-up in a composite map through its bucket directory (compose.cu): the piece of rank r lies between dir[r >> s] and
// dir[(r >> s) + 1], so the binary search runs over a window of a few entries (the widest window of the warp sets the
// trip count).  Only the slots in `act` advance.
template<int C>
__device__ __forceinline__ void lookup_comp(uint32_t (&r)[C], uint32_t tab_saddr, uint32_t td_off, uint32_t dir_saddr, int sh, uint32_t act)
{
	if (!__any_sync(FULL_MASK, act != 0)) return;
	uint32_t a[C], hi[C], w = 1;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t d = dir_saddr + ((r[c] >> sh) << 1);
		const uint32_t lo = lds_u16(d), h = lds_u16(d + 2);
		a[c] = tab_saddr + (lo << 2); hi[c] = tab_saddr + (h << 2);
		const uint32_t wc = h - lo + 1u;
		w = wc > w ? wc : w;
	}
	w = __reduce_max_sync(FULL_MASK, w);
	for (uint32_t len = w; len > 1;) {
		const uint32_t half = len >> 1, h4 = half << 2;
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const uint32_t t = a[c] + h4;
			const uint32_t v = t <= hi[c] ? lds_u32(t) : 0xffffffffu;
			a[c] = v <= r[c] ? t : a[c];
		}
		len -= half;
	}
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t d = lds_u32(a[c] + td_off);
		if ((act >> c) & 1u) r[c] += d;
	}
}

// QUERY mode: only the slots in `act` advance (bit c = slot c); a warp in which no lane has slot c active skips its search
template<int C>
__device__ __forceinline__ void lookup_runs_masked(uint32_t (&r)[C], uint32_t ts_saddr, uint32_t n, uint32_t zeros_total, uint32_t act, uint32_t &bits)
{
	if (__all_sync(FULL_MASK, act == (1u << C) - 1u)) { lookup_runs<C>(r, ts_saddr, n, zeros_total, bits); return; }
	uint32_t b = 0;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const bool on = (act >> c) & 1u;
		if (!__any_sync(FULL_MASK, on)) continue;
		uint32_t a = ts_saddr;
		for (uint32_t len = n; len > 1;) {
			const uint32_t half = len >> 1, t = a + (half << 2);
			a = lds_u32(t) <= r[c] ? t : a;
			len -= half;
		}
		const uint32_t nr = r[c] + lds_u32(a + TD_MINUS_TS);
		if (on) { r[c] = nr; b |= (nr >= zeros_total ? 1u : 0u) << c; }
	}
	bits = b;
}

// same, restricted to the columns whose rank lies in [cs, ce) and that were not resolved by an earlier piece
template<int C>
__device__ __forceinline__ void lookup_runs_piece(uint32_t (&r)[C], uint32_t ts_saddr, uint32_t n, uint32_t zeros_total,
                                                  uint32_t cs, uint32_t ce, uint32_t &done, uint32_t &bits)
{
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const bool act = !((done >> c) & 1u) && r[c] >= cs && r[c] < ce;
		if (!act) continue;
		uint32_t a = ts_saddr;
		for (uint32_t len = n; len > 1;) {
			const uint32_t half = len >> 1, t = a + (half << 2);
			a = lds_u32(t) <= r[c] ? t : a;
			len -= half;
		}
		r[c] += lds_u32(a + TD_MINUS_TS);
		bits |= (r[c] >= zeros_total ? 1u : 0u) << c;
		done |= 1u << c;
	}
}

struct WalkSmem {
	uint64_t *mbar;
	uint8_t  *raw;
	uint32_t *ts;
	int32_t  *td;
	RowMeta  *meta;
	int32_t  *rowcnt;   // [T_MAX][G][3]
	uint8_t  *gmask;    // [G][NT]: bit c = slot c of this thread belongs to group g (generic-G path only)
	uint32_t *scratch;  // [8]
};

// composite stages of the QUERY walk (three, so that two TMA copies are in flight while one map is searched): piece starts,
// translations, bucket directory -- overlaid on the raw / run-table buffers, which the composite phase does not use
constexpr uint32_t CSTG_TD = COMP_CAP * 4u, CSTG_DIR = COMP_CAP * 8u, CSTG_BYTES = COMP_CAP * 8u + COMP_DIR * 2u;
constexpr int CSTG_N = 3;

__host__ __device__ inline size_t walk_smem_layout(int C, int G, size_t off[8], bool query = false)
{
	size_t o = 0;
	off[0] = o; o += 32;                                  // mbarriers (tile copy / composite stage 0, stages 1 and 2)
	off[1] = o; o += RAW_BYTES;                           // raw
	off[2] = o; o += sizeof(uint32_t) * RAW_BYTES;        // ts
	off[3] = o; o += sizeof(int32_t) * RAW_BYTES;         // td (must directly follow ts: TD_MINUS_TS)
	if (query && o - off[1] < (size_t)CSTG_N * CSTG_BYTES) o = off[1] + (size_t)CSTG_N * CSTG_BYTES;
	off[4] = o; o += sizeof(RowMeta) * T_MAX;             // meta
	off[5] = o; o += sizeof(int32_t) * T_MAX * G * 3;     // rowcnt
	off[6] = o; o += (G > 2 ? (size_t)G * WALK_NT : 16);  // gmask
	o = (o + 15) & ~(size_t)15;
	off[7] = o; o += sizeof(uint32_t) * 8;                // scratch
	(void)C;
	return (o + 15) & ~(size_t)15;
}

size_t walk_smem_bytes(int C, int G, bool query) { size_t off[8]; return walk_smem_layout(C, G, off, query); }

// Per-row reduction (bgt.c:743-756).  bits0/bits1: bit c = plane-0/1 bit of this thread's column slot c (already
// masked to valid slots).  Every thread popcounts its own C codes per group, one REDUX.SUM per counter folds the
// warp, lane 0 adds to the tile's shared counters.  gm0/gm1: per-thread slot masks of groups 1 and 2 (G <= 2).
template<int C>
__device__ __forceinline__ void count_row(uint32_t bits0, uint32_t bits1, bool zero1, uint32_t gm0, uint32_t gm1,
                                          const WalkSmem &S, int G, int r_in_tile, int tid, int lane)
{
	const uint32_t x1 = bits0 & ~bits1, x2 = ~bits0 & bits1, x3 = bits0 & bits1;  // ALT, missing, other-ALT
	int32_t *rc = S.rowcnt + r_in_tile * G * 3;
	if (G <= 2) {
		#pragma unroll
		for (int g = 0; g < 2; ++g) {
			if (g >= G) break;
			const uint32_t gm = g ? gm1 : gm0;
			const int s1 = __reduce_add_sync(FULL_MASK, __popc(x1 & gm));
			if (lane == 0 && s1) atomicAdd(rc + g * 3, s1);
			if (!zero1) {
				const int s2 = __reduce_add_sync(FULL_MASK, __popc(x2 & gm));
				const int s3 = __reduce_add_sync(FULL_MASK, __popc(x3 & gm));
				if (lane == 0 && s2) atomicAdd(rc + g * 3 + 1, s2);
				if (lane == 0 && s3) atomicAdd(rc + g * 3 + 2, s3);
			}
		}
	} else {
		for (int g = 0; g < G; ++g) {
			const uint32_t gm = S.gmask[g * WALK_NT + tid];
			const int s1 = __reduce_add_sync(FULL_MASK, __popc(x1 & gm));
			if (lane == 0 && s1) atomicAdd(rc + g * 3, s1);
			if (!zero1) {
				const int s2 = __reduce_add_sync(FULL_MASK, __popc(x2 & gm));
				const int s3 = __reduce_add_sync(FULL_MASK, __popc(x3 & gm));
				if (lane == 0 && s2) atomicAdd(rc + g * 3 + 1, s2);
				if (lane == 0 && s3) atomicAdd(rc + g * 3 + 2, s3);
			}
		}
	}
}

// genotype rows as two bit planes: one ballot per 32 consecutive tracked columns per plane
template<int C>
__device__ __forceinline__ void emit_row(uint32_t bits0, uint32_t bits1, const WalkParams &P, long long out_row,
                                         int warp, int lane, int slice_base)
{
	uint32_t w0 = 0, w1 = 0;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const uint32_t b0 = __ballot_sync(FULL_MASK, (bits0 >> c) & 1u);
		const uint32_t b1 = __ballot_sync(FULL_MASK, (bits1 >> c) & 1u);
		if (lane == c) { w0 = b0; w1 = b1; }
	}
	if (lane < C) {
		const int wi = (slice_base + lane * WALK_NT) / 32 + warp;
		if (wi < P.words) {
			P.hap[0][(size_t)out_row * P.words + wi] = w0;
			P.hap[1][(size_t)out_row * P.words + wi] = w1;
		}
	}
}

// MODE: what the walk produces besides advancing the ranks
//   WALK_COUNT  per-site per-group counters                      (the `view -G` scan)
//   WALK_EMIT   counters + genotype rows as bit planes           (seam A, `view` with genotypes, subsets)
//   WALK_CHAIN  nothing per site; one chain through ALL blocks that dumps the running permutation in front of
//               every block (synthetic cohort generator)
//   WALK_QUERY  second phase of the split scan (api.cu): the tracked entries are (column, target row) pairs -- the
//               haplotypes that carry a missing / other-ALT code at that row (plane1_select_kernel).  Only plane 0 is
//               walked, an entry contributes its code at its target row and the CTA stops after its last target.
enum { WALK_COUNT = 0, WALK_EMIT = 1, WALK_CHAIN = 2, WALK_QUERY = 3 };

template<int C, int MODE>
__global__ void __launch_bounds__(WALK_NT, 2) pbwt_walk_kernel(const WalkParams P)
{
	constexpr bool EMIT = MODE == WALK_EMIT, CHAIN = MODE == WALK_CHAIN, QUERY = MODE == WALK_QUERY;
	constexpr bool COUNT = MODE == WALK_COUNT || MODE == WALK_EMIT;
	extern __shared__ __align__(128) uint8_t smem[];
	WalkSmem S;
	{
		size_t off[8];
		walk_smem_layout(C, P.G, off, QUERY);
		S.mbar = (uint64_t*)(smem + off[0]); S.raw = smem + off[1]; S.ts = (uint32_t*)(smem + off[2]); S.td = (int32_t*)(smem + off[3]);
		S.meta = (RowMeta*)(smem + off[4]); S.rowcnt = (int32_t*)(smem + off[5]); S.gmask = smem + off[6];
		S.scratch = (uint32_t*)(smem + off[7]);
	}
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int BS = 1 << P.shift;
	const int slice_base = blockIdx.x * (WALK_NT * C);
	const uint32_t m = (uint32_t)P.m;
	// the checkpoint block of this CTA (a list when only some blocks take this path) and its tracked columns
	// (one list for the whole scan, or one list per block when the set was derived per block)
	const int blk_own = CHAIN ? 0 : (P.blk_list ? P.blk_list[blockIdx.y] : P.blk_first + (int)blockIdx.y);
	if (!CHAIN && ((P.blk_skip && P.blk_skip[blk_own]) || (P.blk_ok && !P.blk_ok[blk_own]))) return;   // another kernel's block (device-side verdict of index.cu)
	const int n_track = P.n_track_blk ? P.n_track_blk[blk_own] : P.n_track;
	const int32_t *track = P.track ? P.track + (size_t)blk_own * P.track_stride : nullptr;
	if (slice_base >= n_track) return;

	// ---- which columns this thread owns, their groups, their start ranks
	uint32_t r0[C], r1[C];
	int32_t col[C];
	uint32_t tgt[C];     // QUERY: target row (within the block) of every entry
	uint32_t validbits = 0, gm0 = 0, gm1 = 0;
	if (P.G > 2) for (int g = 0; g < P.G; ++g) S.gmask[g * WALK_NT + tid] = 0;
	#pragma unroll
	for (int c = 0; c < C; ++c) {
		const int e = slice_base + c * WALK_NT + tid;
		const bool v = e < n_track;
		col[c] = v ? (track ? track[e] : e) : 0;
		validbits |= (v ? 1u : 0u) << c;
		tgt[c] = (QUERY && v) ? (uint32_t)P.qrow[(size_t)blk_own * P.track_stride + e] : 0xffffffffu;
		if (v && (COUNT || QUERY)) {
			const int grp = (int)P.tgrp[P.track_stride ? col[c] : e];   // per-block lists: groups are per column
			if (grp == 0) gm0 |= 1u << c;
			else if (grp == 1) gm1 |= 1u << c;
			if (P.G > 2) S.gmask[grp * WALK_NT + tid] |= (uint8_t)(1u << c);
		}
		r0[c] = r1[c] = CHAIN ? (uint32_t)col[c] : 0u; // generator: identity before row 0 (pbwt.c:103)
	}
	for (int i = tid; i < T_MAX * P.G * 3; i += WALK_NT) S.rowcnt[i] = 0;
	if (tid == 0) { mbar_init(S.mbar, 1); mbar_init(S.mbar + 1, 1); mbar_init(S.mbar + 2, 1); }
	__syncthreads();

	uint32_t parity = 0;
	const uint32_t ts_saddr = smem_u32(S.ts);
	const int blk_lo = blk_own;
	const int blk_hi = CHAIN ? P.n_blk_chain : blk_lo + 1;
	// QUERY: entries are sorted by target row, so the last valid entry of this slice bounds the rows this CTA needs
	int q_stop = 0x7fffffff;
	if (QUERY) {
		const int last = (slice_base + WALK_NT * C < n_track ? slice_base + WALK_NT * C : n_track) - 1;
		q_stop = (int)P.qrow[(size_t)blk_own * P.track_stride + last] + 1;
	}

	for (int blk = blk_lo; blk < blk_hi; ++blk) {
		const long long blk_row = P.blk_row0 + ((long long)blk << P.shift);
		const long long row_hi = (QUERY && blk_row + q_stop < P.row_hi) ? blk_row + q_stop : P.row_hi;
		const uint64_t *roff = P.rowoff + (size_t)blk * (BS + 1);
		if (CHAIN) { // pbwt.c:292-301: dump the running permutation of both planes in front of the block
			uint8_t *Sp = P.snap_img + P.blkoff[blk] + 1;
			#pragma unroll
			for (int c = 0; c < C; ++c)
				if ((validbits >> c) & 1u) {
					uint8_t *q0 = Sp + 4 * (size_t)r0[c], *q1 = Sp + 4 * (size_t)m + 4 * (size_t)r1[c];
					const uint32_t v = (uint32_t)col[c];
					q0[0] = v; q0[1] = v >> 8; q0[2] = v >> 16; q0[3] = v >> 24;
					q1[0] = v; q1[1] = v >> 8; q1[2] = v >> 16; q1[3] = v >> 24;
				}
		} else {
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				r0[c] = (uint32_t)P.rank0[((size_t)blk * 2 + 0) * m + col[c]];
				r1[c] = (uint32_t)P.rank0[((size_t)blk * 2 + 1) * m + col[c]];
			}
		}
		int t_beg = P.blk_tile_beg[blk];
		const int t_end = P.blk_tile_end[blk];
		uint32_t from_row[C];     // QUERY: first row (within the block) that the entry walks row by row
		#pragma unroll
		for (int c = 0; c < C; ++c) from_row[c] = 0;
		if (QUERY && P.comp_n) {
			// Every entry crosses the row groups that lie entirely in front of ITS target row with ONE look-up per group in
			// the group's composite map (compose.cu), staged by TMA straight into the run-table buffers; it then walks row by
			// row only inside its own group (below).  Entries are sorted by target row, so a group is live for a suffix of
			// the CTA's entries and warps without a live entry skip the search.
			const int n_grp = (BS + COMP_K - 1) / COMP_K;
			const int g_mixed = (int)P.qrow[(size_t)blk_own * P.track_stride + slice_base] / COMP_K;
			const int g_last = (q_stop - 1) / COMP_K;            // group of the CTA's last target row
			// g_avail: the groups in front of the first one without a usable composite (all of them, as a rule)
			__shared__ int s_avail;
			if (tid == 0) s_avail = g_last;
			__syncthreads();
			for (int i = tid; i < g_last; i += WALK_NT) if (P.comp_n[(size_t)blk * n_grp + i] == 0) atomicMin(&s_avail, i);
			__syncthreads();
			const int g_avail = s_avail;
			// three stages: while group g is searched, the copies of g+1 and g+2 are in flight (a TMA round trip is longer
			// than one search); stage s has its own mbarrier and is re-filled after the CTA barrier that ends its last use
			uint8_t *stg0 = S.raw;
			const uint32_t stg0_saddr = smem_u32(stg0);
			auto fetch_comp = [&](int gg) {
				const size_t slot = (size_t)blk * n_grp + gg;
				const uint32_t np = (uint32_t)P.comp_n[slot];
				uint64_t *bar = S.mbar + gg % CSTG_N;
				uint8_t *dst = stg0 + (size_t)(gg % CSTG_N) * CSTG_BYTES;
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				mbar_expect_tx(bar, np * 8u + (uint32_t)P.dir_n * 2u);
				tma_bulk_g2s(dst, P.comp_start + slot * COMP_CAP, np * 4u, bar);
				tma_bulk_g2s(dst + CSTG_TD, P.comp_delta + slot * COMP_CAP, np * 4u, bar);
				tma_bulk_g2s(dst + CSTG_DIR, P.comp_dir + slot * COMP_DIR_STRIDE, (uint32_t)P.dir_n * 2u, bar);
			};
			int g = 0;
			if (tid == 0) for (int k = 0; k < CSTG_N - 1 && k < g_avail; ++k) fetch_comp(k);
			for (; g < g_avail; ++g) {
				if (tid == 0 && g + CSTG_N - 1 < g_avail) fetch_comp(g + CSTG_N - 1);   // its stage was released by the barrier below
				{
					uint32_t spins = 0;
					const uint32_t par = (uint32_t)(g / CSTG_N) & 1u;
					while (!mbar_try_wait(S.mbar + g % CSTG_N, par))
						if (++spins > (1u << 26)) { atomicOr(P.err, 8); __trap(); }
				}
				uint32_t act = 0;
				#pragma unroll
				for (int c = 0; c < C; ++c) act |= (tgt[c] != 0xffffffffu && (int)(tgt[c] / COMP_K) > g ? 1u : 0u) << c;
				const uint32_t tab = stg0_saddr + (uint32_t)(g % CSTG_N) * CSTG_BYTES;
				lookup_comp<C>(r0, tab, CSTG_TD, tab + CSTG_DIR, P.dir_shift, act);
				__syncthreads();
			}
			parity = (uint32_t)((g_avail + CSTG_N - 1) / CSTG_N) & 1u;   // phases completed on barrier 0, which the tile loop uses next
			// g = first group without a usable composite (or the last group): entries with a later target walk from there
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const uint32_t tg = tgt[c] / COMP_K;
				from_row[c] = (tg < (uint32_t)g ? tg : (uint32_t)g) * COMP_K;
			}
			t_beg = P.grp_tile_beg[(size_t)blk * (n_grp + 1) + (g_mixed < g ? g_mixed : g)];
		}

		// QUERY: the rows at which some entry of this warp is live, per slot (warp-uniform; entries are sorted by target row,
		// so the 32 entries of a slot are neighbours) -- lets the row loop skip dead rows before touching anything
		uint32_t wlo[C], whi[C];
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const bool v = QUERY && tgt[c] != 0xffffffffu;
			wlo[c] = QUERY ? __reduce_min_sync(FULL_MASK, v ? from_row[c] : 0xffffffffu) : 0u;
			whi[c] = QUERY ? __reduce_max_sync(FULL_MASK, v ? tgt[c] : 0u) : 0xffffffffu;
		}

		// thread 0: start the TMA bulk copy of tile t if it is an ordinary (not oversized) tile that will be used
		auto prefetch = [&](int t) {
			if (t >= t_end) return;
			const int2 tl = P.tiles[t];
			if (tl.y < 0 || blk_row + tl.x >= row_hi) return;
			const uint64_t beg = roff[tl.x] & ~(uint64_t)15, end = (roff[tl.x + tl.y] + 15) & ~(uint64_t)15;
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic-proxy accesses to raw[] before the async-proxy write
			mbar_expect_tx(S.mbar, (uint32_t)(end - beg));
			tma_bulk_g2s(S.raw, P.img + beg, (uint32_t)(end - beg), S.mbar);
		};
		if (tid == 0) prefetch(t_beg);

		for (int t = t_beg; t < t_end; ++t) {
			const int2 tl = P.tiles[t];
			const int r_first = tl.x, nr = tl.y & 0x7fffffff;
			const bool big = tl.y < 0;
			if (blk_row + r_first >= row_hi) break;

			if (!big) {
				// ---- wait for the tile's bytes
				{
					uint32_t spins = 0;
					while (!mbar_try_wait(S.mbar, parity))
						if (++spins > (1u << 26)) { atomicOr(P.err, 8); __trap(); }
					parity ^= 1;
				}
				const uint64_t src_beg = roff[r_first] & ~(uint64_t)15;
				if (tid < nr) {
					RowMeta mt;
					const uint32_t o = (uint32_t)(roff[r_first + tid] - src_beg); // 'B'
					mt.len[0] = lds_u32_bytes(S.raw + o + 1);
					mt.off[0] = o + 5;
					mt.len[1] = lds_u32_bytes(S.raw + o + 5 + mt.len[0]);
					mt.off[1] = o + 9 + mt.len[0];
					const uint32_t *n1p = P.n1 + ((size_t)blk * BS + r_first + tid) * 2;
					mt.n1[0] = n1p[0]; mt.n1[1] = n1p[1];
					S.meta[tid] = mt;
				}
				__syncthreads();
				// ---- parse: one warp per (row, plane); constant planes need no table
				for (int task = warp; task < nr * 2; task += WALK_NW) {
					const int r = task >> 1, p = task & 1;
					if (QUERY && p == 1) continue;
					const RowMeta &mt = S.meta[r];
					const uint32_t n1 = mt.n1[p];
					if (n1 == 0 || n1 == m) continue;
					uint32_t tot = 0, ones = 0;
					parse_runs(S.raw, mt.off[p], mt.len[p], m - n1, S.ts, S.td, lane, tot, ones);
				}
				__syncthreads();
				if (tid == 0) prefetch(t + 1); // raw bytes are dead now: overlap the next copy with the walk
				// ---- walk
				for (int r = 0; r < nr; ++r) {
					const long long arow = blk_row + r_first + r;
					if (arow >= row_hi) break;
					if (QUERY) { // warp-uniform: no entry of this warp is live at this row
						const uint32_t rb = (uint32_t)(r_first + r);
						bool live = false;
						#pragma unroll
						for (int c = 0; c < C; ++c) live = live || (rb >= wlo[c] && rb <= whi[c]);
						if (!live) continue;
					}
					const RowMeta mt = S.meta[r];
					uint32_t bits0 = 0, bits1 = 0;
					const bool triv0 = mt.n1[0] == 0 || mt.n1[0] == m, triv1 = mt.n1[1] == 0 || mt.n1[1] == m;
					if (QUERY) {
						// an entry is live from the start of its own group (or of the first group without a composite) to its target row
						const uint32_t rb = (uint32_t)(r_first + r);
						uint32_t act = 0;
						#pragma unroll
						for (int c = 0; c < C; ++c) act |= (rb >= from_row[c] && rb <= tgt[c] && tgt[c] != 0xffffffffu ? 1u : 0u) << c;
						if (!triv0) lookup_runs_masked<C>(r0, ts_saddr + 4u * mt.off[0], mt.len[0], m - mt.n1[0], act, bits0);
						else if (mt.n1[0]) bits0 = 0xffffffffu;
					} else if (!triv0) lookup_runs<C>(r0, ts_saddr + 4u * mt.off[0], mt.len[0], m - mt.n1[0], bits0);
					else if (mt.n1[0]) bits0 = 0xffffffffu;
					if (!QUERY) {
						if (!triv1) lookup_runs<C>(r1, ts_saddr + 4u * mt.off[1], mt.len[1], m - mt.n1[1], bits1);
						else if (mt.n1[1]) bits1 = 0xffffffffu;
					}
					if (COUNT && arow >= P.row_lo) {
						const bool zero0 = mt.n1[0] == 0, zero1 = mt.n1[1] == 0;
						if (!(zero0 && zero1))
							count_row<C>(bits0 & validbits, bits1 & validbits, zero1, gm0, gm1, S, P.G, r, tid, lane);
						if (EMIT) emit_row<C>(bits0 & validbits, bits1 & validbits, P, arow - P.row_lo, warp, lane, slice_base);
					}
					if (QUERY && arow >= P.row_lo) { // entries whose target is this row: code 3 if their plane-0 bit is set, else 2
						uint32_t hit = 0;
						#pragma unroll
						for (int c = 0; c < C; ++c) hit |= (tgt[c] == (uint32_t)(r_first + r) ? 1u : 0u) << c;
						if (__any_sync(FULL_MASK, hit != 0)) count_row<C>(bits0 & hit, hit, false, gm0, gm1, S, P.G, r, tid, lane);
					}
				}
			} else {
				// ---- a single row larger than the staging buffer: stream it in pieces of RAW_CAP bytes
				const long long arow = blk_row + r_first;
				const uint8_t *rec = P.img + roff[r_first];
				const uint32_t *n1p = P.n1 + ((size_t)blk * BS + r_first) * 2;
				uint32_t bits[2] = {0, 0};
				const uint8_t *pp = rec + 1;
				for (int p = 0; p < 2; ++p) {
					const uint32_t l = ld_u32_unaligned(pp), n1 = n1p[p];
					const uint8_t *rle = pp + 4;
					pp = rle + l;
					if (QUERY && p == 1) continue;
					if (n1 == 0 || n1 == m) { bits[p] = n1 ? 0xffffffffu : 0u; continue; }
					uint32_t done = 0;
					if (QUERY) { // entries that are not live at this row (see the tile path) must not move
						#pragma unroll
						for (int c = 0; c < C; ++c)
							done |= ((uint32_t)r_first >= from_row[c] && (uint32_t)r_first <= tgt[c] && tgt[c] != 0xffffffffu ? 0u : 1u) << c;
					}
					if (tid == 0) { S.scratch[0] = 0; S.scratch[1] = 0; }
					for (uint32_t cb = 0; cb < l; cb += RAW_CAP) {
						const uint32_t n = l - cb < (uint32_t)RAW_CAP ? l - cb : (uint32_t)RAW_CAP;
						for (uint32_t i = tid; i < n; i += WALK_NT) S.raw[i] = rle[cb + i];
						__syncthreads();
						const uint32_t cs = S.scratch[0];
						if (warp == 0) {
							uint32_t tot = cs, ones = S.scratch[1];
							parse_runs(S.raw, 0, n, m - n1, S.ts, S.td, lane, tot, ones);
							if (lane == 0) { S.scratch[2] = tot; S.scratch[3] = ones; }
						}
						__syncthreads();
						const uint32_t ce = S.scratch[2];
						if (p == 0) lookup_runs_piece<C>(r0, ts_saddr, n, m - n1, cs, ce, done, bits[0]);
						else        lookup_runs_piece<C>(r1, ts_saddr, n, m - n1, cs, ce, done, bits[1]);
						__syncthreads();
						if (tid == 0) { S.scratch[0] = S.scratch[2]; S.scratch[1] = S.scratch[3]; }
					}
					__syncthreads();
				}
				if (COUNT && arow >= P.row_lo && arow < row_hi) {
					if (n1p[0] || n1p[1])
						count_row<C>(bits[0] & validbits, bits[1] & validbits, n1p[1] == 0, gm0, gm1, S, P.G, 0, tid, lane);
					if (EMIT) emit_row<C>(bits[0] & validbits, bits[1] & validbits, P, arow - P.row_lo, warp, lane, slice_base);
				}
				if (QUERY && arow >= P.row_lo && arow < row_hi) {
					uint32_t hit = 0;
					#pragma unroll
					for (int c = 0; c < C; ++c) hit |= (tgt[c] == (uint32_t)r_first ? 1u : 0u) << c;
					if (__any_sync(FULL_MASK, hit != 0)) count_row<C>(bits[0] & hit, hit, false, gm0, gm1, S, P.G, 0, tid, lane);
				}
				if (tid == 0) prefetch(t + 1);
			}

			// ---- flush the tile's counters
			__syncthreads();
			if (COUNT || QUERY) {
				const int per_row = P.G * 3;
				for (int i = tid; i < nr * per_row; i += WALK_NT) {
					const int v = S.rowcnt[i];
					if (v) {
						const long long arow = blk_row + r_first + i / per_row;
						if (arow >= P.row_lo && arow < row_hi)
							atomicAdd(P.cnt_raw + (size_t)(arow - P.row_lo) * per_row + i % per_row, v);
						S.rowcnt[i] = 0;
					}
				}
			}
		}
		if (CHAIN) __syncthreads();
	}
}

template<int C, int MODE>
static cudaError_t launch_walk_t(const WalkParams &P, int slices, int n_blk, cudaStream_t st)
{
	const size_t smem = walk_smem_bytes(C, P.G, MODE == WALK_QUERY);
	cudaError_t e = cudaFuncSetAttribute(pbwt_walk_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	for (int b0 = 0; b0 < n_blk; b0 += 32768) {
		WalkParams Q = P;
		if (P.blk_list) Q.blk_list = P.blk_list + b0; else Q.blk_first = P.blk_first + b0;
		const int nb = n_blk - b0 < 32768 ? n_blk - b0 : 32768;
		dim3 grid(slices, MODE == WALK_CHAIN ? 1 : nb, 1);
		pbwt_walk_kernel<C, MODE><<<grid, WALK_NT, smem, st>>>(Q);
		if (MODE == WALK_CHAIN) break;
	}
	return cudaGetLastError();
}

template<int MODE>
static cudaError_t launch_walk_c(const WalkParams &P, int C, int slices, int n_blk, cudaStream_t st)
{
	switch (C) {
	case 1: return launch_walk_t<1, MODE>(P, slices, n_blk, st);
	case 2: return launch_walk_t<2, MODE>(P, slices, n_blk, st);
	case 4: return launch_walk_t<4, MODE>(P, slices, n_blk, st);
	default: return launch_walk_t<8, MODE>(P, slices, n_blk, st);
	}
}

cudaError_t launch_walk(const WalkParams &P, int C, int mode, int slices, int n_blk, cudaStream_t st)
{
	if (slices <= 0 || n_blk <= 0) return cudaSuccess;
	switch (mode) {
	case WALK_EMIT: return launch_walk_c<WALK_EMIT>(P, C, slices, n_blk, st);
	case WALK_CHAIN: return launch_walk_c<WALK_CHAIN>(P, C, slices, n_blk, st);
	case WALK_QUERY: return launch_walk_c<WALK_QUERY>(P, C, slices, n_blk, st);
	default: return launch_walk_c<WALK_COUNT>(P, C, slices, n_blk, st);
	}
}

// ------------------------------------------------------------------------------------------------ finalize

// Per site: (#ALT, #missing, #other-ALT) per group -> bgt_info_t (bgt.c:745-756) -> filter verdict (bgt.c:712-719).
__global__ void __launch_bounds__(256) finalize_kernel(const int32_t *__restrict__ cnt_raw, long long n_rows, int G,
                                                       const int32_t *__restrict__ gsize, const flt_prog_t *__restrict__ prog, int use_flt,
                                                       int32_t *__restrict__ counts, uint8_t *__restrict__ pass,
                                                       unsigned long long *__restrict__ totals, const FinalizeSplit sp)
{
	const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long t_an = 0, t_ac0 = 0, t_ac1 = 0, t_pass = 0;
	if (row < n_rows) {
		int32_t v[3 + 3 * B200_MAX_GROUPS_K];
		const int32_t *c = cnt_raw + (size_t)row * G * 3;
		int32_t an = 0, ac0 = 0, ac1 = 0;
		for (int g = 0; g < G; ++g) {
			int32_t c1 = c[g * 3];
			const int32_t c2 = c[g * 3 + 1], c3 = c[g * 3 + 2];
			if (sp.blk_split) { // split scan: the walk only counted the plane-1 codes; #ALT = (plane-0 ones in the group) - #other-ALT
				const long long arow = sp.row_lo + row;
				const int blk = (int)((arow - sp.blk_row0) >> sp.shift);
				if (sp.blk_split[blk]) {
					int32_t n0;
					if (g < sp.n_vec) n0 = sp.n0g[(size_t)row * sp.n_vec + g];
					else {
						n0 = (int32_t)sp.n1[((size_t)blk << sp.shift << 1) + (size_t)(arow - sp.blk_row0 - ((long long)blk << sp.shift)) * 2];
						for (int k = 0; k < sp.n_vec; ++k) n0 -= sp.n0g[(size_t)row * sp.n_vec + k];
					}
					c1 = n0 - c3;
				}
			}
			const int32_t gan = gsize[g] - c2;            // c0 + c1 + c3
			v[3 + 3 * g] = gan; v[4 + 3 * g] = c1; v[5 + 3 * g] = c3;
			an += gan; ac0 += c1; ac1 += c3;
		}
		v[0] = an; v[1] = ac0; v[2] = ac1;
		const int stride = 3 + 3 * G;
		if (counts) for (int k = 0; k < stride; ++k) counts[(size_t)row * stride + k] = v[k];
		const int ok = use_flt ? flt_eval(prog, v) : 1;
		if (pass) pass[row] = (uint8_t)ok;
		t_an = an; t_ac0 = ac0; t_ac1 = ac1; t_pass = ok;
	}
	#pragma unroll
	for (int d = 16; d; d >>= 1) {
		t_an += __shfl_xor_sync(FULL_MASK, t_an, d); t_ac0 += __shfl_xor_sync(FULL_MASK, t_ac0, d);
		t_ac1 += __shfl_xor_sync(FULL_MASK, t_ac1, d); t_pass += __shfl_xor_sync(FULL_MASK, t_pass, d);
	}
	if ((threadIdx.x & 31) == 0 && totals) {
		if (t_an) atomicAdd(totals + 0, (unsigned long long)t_an);
		if (t_ac0) atomicAdd(totals + 1, (unsigned long long)t_ac0);
		if (t_ac1) atomicAdd(totals + 2, (unsigned long long)t_ac1);
		if (t_pass) atomicAdd(totals + 3, (unsigned long long)t_pass);
	}
}

cudaError_t launch_finalize(const int32_t *cnt_raw, long long n_rows, int G, const int32_t *gsize, const flt_prog_t *prog, int use_flt,
                            int32_t *counts, uint8_t *pass, unsigned long long *totals, const FinalizeSplit &sp, cudaStream_t st)
{
	if (n_rows <= 0) return cudaSuccess;
	finalize_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(cnt_raw, n_rows, G, gsize, prog, use_flt, counts, pass, totals, sp);
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ bit planes -> bytes

// pbf_read's layout (pbwt.c:313-337 returns one byte per haplotype): thread = 4 haplotypes -> one 32-bit store
__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint32_t *__restrict__ bits, long long n_rows, int words, int n_track,
                                                          uint8_t *__restrict__ bytes)
{
	const int quads = (n_track + 3) / 4;
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rows * quads) return;
	const long long row = i / quads;
	const int q = (int)(i % quads), h = q * 4;
	const uint32_t w = bits[(size_t)row * words + (h >> 5)] >> (h & 31);
	uint8_t *dst = bytes + (size_t)row * n_track + h;
	#pragma unroll
	for (int k = 0; k < 4; ++k) if (h + k < n_track) dst[k] = (w >> k) & 1u;
}

cudaError_t launch_unpack_bits(const uint32_t *bits, long long n_rows, int words, int n_track, uint8_t *bytes, cudaStream_t st)
{
	const long long n = n_rows * ((n_track + 3) / 4);
	if (n <= 0) return cudaSuccess;
	unpack_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bits, n_rows, words, n_track, bytes);
	return cudaGetLastError();
}

} // namespace b200
