// pairwalk.cuh -- device helpers shared by the two "pair" kernels of the split scan: pairwalk.cu (plane-0 bit of every
// (column,row) pair) and plane1.cu (which columns carry the plane-1 bits).  Both move a sorted list of ranks through row
// groups by composite maps staged in shared memory (TMA bulk copies + mbarriers) and through single rows by run tables a warp
// builds for itself.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

#define PW_FULL 0xffffffffu

constexpr int PW_NT = PAIR_SLICE_THREADS, PW_NW = PW_NT / 32;
constexpr int PW_STAGES = 3;
constexpr uint32_t PW_STG_TD = COMP_CAP * 4u, PW_STG_DIR = COMP_CAP * 8u, PW_STG_BYTES = COMP_CAP * 8u + COMP_DIR_STRIDE * 2u;
// phase B: every warp owns a slice of the (then idle) stages: a run table of PW_TAB2 entries (ts + td) and PW_RAW staged record bytes
constexpr int PW_SLICE = (int)((PW_STAGES * PW_STG_BYTES) / PW_NW) & ~15;
constexpr int PW_TAB2 = 256;
constexpr int PW_RAW = PW_SLICE - PW_TAB2 * 8;

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
template<int C, uint32_t TD_OFF = PW_STG_TD>
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
		const uint32_t d = pw_lds_u32(a[c] + TD_OFF);
		if ((act >> c) & 1u) r[c] += d;
	}
}

// One warp, one row.  The row's RLE bytes (generic pointer: the warp's staged copy in shared memory, or global memory for a
// record larger than the slice) are turned into a run table in the warp's slice, PW_TAB2 entries a piece: four bytes per lane
// and pass -- a local prefix over the lane's bytes, one warp scan over the lane totals -- so a typical row (100-200 bytes)
// takes two passes.  Then every slot c whose bit is set in the lane's `act` moves its rank through the row.
//   BACK = false: rank in front of the row -> rank behind it (rank' = rank + delta of the run that holds it, pbwt.c:150);
//                 `bits` bit c = the run's bit.  Table: one entry per byte in file order (ts = run start, td = delta).
//   BACK = true : rank behind the row -> rank in front of it (the partition of pbwt.c:79-88 undone).  A rank below the zeros
//                 total was the k-th 0 of the row, one above it the k-th 1.  Table, again per byte in file order: zb = zeros
//                 in front of the byte, ob = ones in front of it -- both ascending, so the k-th 0 lies in the last byte with
//                 zb <= k (bytes of the other class and empty bytes share the zb of the 0-byte behind them and are never last),
//                 at rank k + ob; the k-th 1 in the last byte with ob <= k, at rank k + zb.
template<int C, bool BACK>
__device__ __forceinline__ void pw_row(const uint8_t *rle, uint32_t l, uint32_t m, uint32_t n1, uint32_t *ts, int lane,
                                       uint32_t (&r)[C], uint32_t act, uint32_t &bits)
{
	const uint32_t zeros_total = m - n1;
	const uint32_t tsa = pw_smem_u32(ts);
	uint32_t tot = 0, ones = 0, todo = act;
	for (uint32_t cb = 0; cb < l; cb += PW_TAB2) {
		const uint32_t n = l - cb < (uint32_t)PW_TAB2 ? l - cb : (uint32_t)PW_TAB2;
		const uint32_t cs = tot, os = ones;
		for (uint32_t base = 0; base < n; base += 128) {
			const uint32_t i0 = base + 4u * lane;
			uint32_t L[4], O[4];
			#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint32_t c = i0 + k < n ? rle[cb + i0 + k] : 0u;
				L[k] = pw_rle_len(c); O[k] = (c & 1u) ? L[k] : 0u;
			}
			const uint32_t xl = L[0] + L[1] + L[2] + L[3], yl = O[0] + O[1] + O[2] + O[3];
			uint32_t x = xl, y = yl;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t tx = __shfl_up_sync(PW_FULL, x, d), ty = __shfl_up_sync(PW_FULL, y, d);
				if (lane >= d) { x += tx; y += ty; }
			}
			uint32_t start = tot + x - xl, ones_before = ones + y - yl;
			if (!BACK) {
				uint4 vs; int4 vd;
				uint32_t *ps = &vs.x; int32_t *pd = &vd.x;
				#pragma unroll
				for (int k = 0; k < 4; ++k) {
					ps[k] = start;
					pd[k] = O[k] ? (int32_t)(zeros_total - (start - ones_before)) : -(int32_t)ones_before;   // (an empty byte is never the entry that is hit)
					start += L[k]; ones_before += O[k];
				}
				*(uint4*)(ts + i0) = vs;
				*(int4*)(ts + PW_TAB2 + i0) = vd;
			} else {
				uint4 vz, vo;
				uint32_t *pz = &vz.x, *po = &vo.x;
				#pragma unroll
				for (int k = 0; k < 4; ++k) {
					pz[k] = start - ones_before; po[k] = ones_before;
					start += L[k]; ones_before += O[k];
				}
				*(uint4*)(ts + i0) = vz;
				*(uint4*)(ts + PW_TAB2 + i0) = vo;
			}
			tot += __shfl_sync(PW_FULL, x, 31);
			ones += __shfl_sync(PW_FULL, y, 31);
		}
		__syncwarp();
		if (!BACK) {
			// all C searches of the lane in lock step (independent shared-memory reads): last entry whose start <= rank -- empty
			// bytes share the start of their successor and are never last.  Slots that are not live, or whose rank lies in
			// another piece of a long row, search along and discard the result.
			uint32_t a[C];
			#pragma unroll
			for (int c = 0; c < C; ++c) a[c] = tsa;
			for (uint32_t len = n; len > 1;) {
				const uint32_t half = len >> 1, h4 = half << 2;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					const uint32_t t = a[c] + h4;
					a[c] = pw_lds_u32(t) <= r[c] ? t : a[c];
				}
				len -= half;
			}
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const uint32_t nr = r[c] + pw_lds_u32(a[c] + PW_TAB2 * 4u);
				if (((todo >> c) & 1u) && r[c] >= cs && r[c] < tot) {
					r[c] = nr;
					bits |= (nr >= zeros_total ? 1u : 0u) << c;
					todo &= ~(1u << c);
				}
			}
		} else {
			// this piece holds the zeros numbered [cs - os, tot - ones) and the ones numbered [os, ones)
			const uint32_t z_lo = cs - os, z_hi = tot - ones;
			uint32_t a[C], kk[C], inp = 0;
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const bool one = r[c] >= zeros_total;
				kk[c] = one ? r[c] - zeros_total : r[c];
				const bool here = one ? (kk[c] >= os && kk[c] < ones) : (kk[c] >= z_lo && kk[c] < z_hi);
				inp |= (((todo >> c) & 1u) && here ? 1u : 0u) << c;
				a[c] = one ? tsa + PW_TAB2 * 4u : tsa;               // search ob[] or zb[]
			}
			for (uint32_t len = n; len > 1;) {
				const uint32_t half = len >> 1, h4 = half << 2;
				#pragma unroll
				for (int c = 0; c < C; ++c) {
					const uint32_t t = a[c] + h4;
					a[c] = pw_lds_u32(t) <= kk[c] ? t : a[c];
				}
				len -= half;
			}
			#pragma unroll
			for (int c = 0; c < C; ++c) {
				const bool one = r[c] >= zeros_total;
				const uint32_t other = pw_lds_u32(one ? a[c] - PW_TAB2 * 4u : a[c] + PW_TAB2 * 4u);   // zb of the byte for a 1, ob for a 0
				if ((inp >> c) & 1u) { r[c] = kk[c] + other; todo &= ~(1u << c); }
			}
		}
		__syncwarp();
		if (!__any_sync(PW_FULL, todo != 0)) break;
	}
}

} // namespace b200
