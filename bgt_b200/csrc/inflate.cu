// inflate.cu -- BGZF block inflate on the device (bgzf.c:225-249 inflate_block, :318-351 bgzf_read_block).
//
// A BGZF file (the site-only .bcf and its .csi index of a BGT database) is a series of gzip members of at most 64 KiB,
// each an 18-byte header with the 'BC' extra field (block size - 1), one raw DEFLATE stream (RFC 1951) and an 8-byte
// trailer CRC32 | ISIZE (bgzf.c:68,76,251-257).  Blocks are independent: one CTA (one warp) per block.  The compressed
// bytes are staged in shared memory by the warp, lane 0 walks the bit stream (canonical Huffman decoding by code
// length, RFC 1951 3.2.2) into a 64 KiB shared-memory window, LZ77 copies are done by the whole warp, and the window is
// written to HBM with coalesced stores.  Like the reference's reader, the CRC is not checked.
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

constexpr int INF_MAXBITS = 15, INF_MAXL = 288, INF_MAXD = 30, INF_WIN = 65536;

struct Huff { uint16_t count[INF_MAXBITS + 1]; uint16_t symbol[INF_MAXL]; };

struct BitIn {
	const uint8_t *p; uint32_t n, pos; uint64_t buf; int cnt; bool over;
	__device__ __forceinline__ void fill() { while (cnt <= 56 && pos < n) { buf |= (uint64_t)p[pos++] << cnt; cnt += 8; } }
	__device__ __forceinline__ uint32_t bits(int k)
	{
		if (k == 0) return 0;
		if (cnt < k) { fill(); if (cnt < k) { over = true; return 0; } }
		const uint32_t v = (uint32_t)(buf & ((1ull << k) - 1));
		buf >>= k; cnt -= k;
		return v;
	}
};

// canonical code from code lengths (RFC 1951 3.2.2); returns <0 for an over-subscribed set, >0 incomplete, 0 complete
__device__ int huff_build(Huff &h, const uint8_t *len, int n)
{
	uint16_t offs[INF_MAXBITS + 1];
	for (int l = 0; l <= INF_MAXBITS; ++l) h.count[l] = 0;
	for (int s = 0; s < n; ++s) ++h.count[len[s]];
	if (h.count[0] == n) return 0;
	int left = 1;
	for (int l = 1; l <= INF_MAXBITS; ++l) { left <<= 1; left -= h.count[l]; if (left < 0) return left; }
	offs[1] = 0;
	for (int l = 1; l < INF_MAXBITS; ++l) offs[l + 1] = offs[l] + h.count[l];
	for (int s = 0; s < n; ++s) if (len[s]) h.symbol[offs[len[s]]++] = (uint16_t)s;
	return left;
}

__device__ __forceinline__ int huff_decode(BitIn &in, const Huff &h)
{
	int code = 0, first = 0, index = 0;
	if (in.cnt < INF_MAXBITS) in.fill();
	for (int l = 1; l <= INF_MAXBITS; ++l) {
		if (in.cnt < 1) { in.over = true; return -1; }
		code |= (int)(in.buf & 1); in.buf >>= 1; --in.cnt;
		const int c = h.count[l];
		if (code - c < first) return h.symbol[index + (code - first)];
		index += c; first += c; first <<= 1; code <<= 1;
	}
	return -1;
}

__constant__ uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t  c_lext[29]  = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t  c_dext[30]  = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t  c_clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// one warp per BGZF block
__global__ void __launch_bounds__(32) bgzf_inflate_kernel(const InflateParams P)
{
	extern __shared__ __align__(16) uint8_t inf_sm[];
	uint8_t *win = inf_sm;                         // [INF_WIN] output window
	uint8_t *cin = inf_sm + INF_WIN;               // [P.max_csize + 16] compressed bytes
	__shared__ Huff hl, hd;
	__shared__ uint8_t lens[INF_MAXL + INF_MAXD + 8];
	__shared__ int s_op, s_a, s_b;                 // lane 0 -> warp: 1 = copy (dist, len) at out position, 2 = done, 3 = error
	__shared__ uint32_t s_out;
	const int lane = threadIdx.x;
	const int b = (int)blockIdx.x;
	const uint64_t coff = P.blk_coff[b];
	const uint32_t csize = P.blk_csize[b], usize = P.blk_usize[b];
	const uint64_t uoff = P.blk_uoff[b];
	// ---- stage the DEFLATE stream (block minus 18-byte header and 8-byte trailer)
	for (uint32_t i = lane; i < csize; i += 32) cin[i] = P.in[coff + i];
	__syncwarp();
	int status = 0;                                // lane 0: 0 running, 2 done, 3 error
	BitIn in;
	uint32_t out = 0;
	if (lane == 0) { in.p = cin; in.n = csize; in.pos = 0; in.buf = 0; in.cnt = 0; in.over = false; }
	bool last = false, in_block = false;
	int btype = 0;
	// The warp runs in lock step: lane 0 decodes until it has a match to copy (or finishes), then all lanes do the copy.
	while (true) {
		if (lane == 0) {
			int op = 0;
			while (op == 0) {
				if (!in_block) {
					if (last) { op = 2; break; }
					last = in.bits(1) != 0;
					btype = (int)in.bits(2);
					if (in.over) { op = 3; break; }
					if (btype == 0) { // stored: to the byte boundary, LEN, NLEN, bytes
						in.buf >>= (in.cnt & 7); in.cnt -= (in.cnt & 7);
						const uint32_t len = in.bits(16), nlen = in.bits(16);
						if (in.over || (len ^ 0xffffu) != nlen || out + len > usize) { op = 3; break; }
						for (uint32_t i = 0; i < len; ++i) { const uint32_t v = in.bits(8); win[out++] = (uint8_t)v; }
						if (in.over) { op = 3; break; }
						continue;
					} else if (btype == 1) { // fixed codes (RFC 1951 3.2.6)
						int s = 0;
						for (; s < 144; ++s) lens[s] = 8;
						for (; s < 256; ++s) lens[s] = 9;
						for (; s < 280; ++s) lens[s] = 7;
						for (; s < 288; ++s) lens[s] = 8;
						huff_build(hl, lens, 288);
						for (s = 0; s < 30; ++s) lens[s] = 5;
						huff_build(hd, lens, 30);
					} else if (btype == 2) { // dynamic codes (RFC 1951 3.2.7)
						const int nlen = (int)in.bits(5) + 257, ndist = (int)in.bits(5) + 1, ncode = (int)in.bits(4) + 4;
						if (in.over || nlen > 286 || ndist > 30) { op = 3; break; }
						int i = 0;
						for (; i < ncode; ++i) lens[c_clorder[i]] = (uint8_t)in.bits(3);
						for (; i < 19; ++i) lens[c_clorder[i]] = 0;
						if (huff_build(hl, lens, 19) != 0) { op = 3; break; }   // the code-length code must be complete
						i = 0;
						bool bad = false;
						while (i < nlen + ndist) {
							const int sym = huff_decode(in, hl);
							if (sym < 0) { bad = true; break; }
							if (sym < 16) lens[i++] = (uint8_t)sym;
							else {
								int prev = 0, rep;
								if (sym == 16) { if (i == 0) { bad = true; break; } prev = lens[i - 1]; rep = 3 + (int)in.bits(2); }
								else if (sym == 17) rep = 3 + (int)in.bits(3);
								else rep = 11 + (int)in.bits(7);
								if (i + rep > nlen + ndist) { bad = true; break; }
								while (rep--) lens[i++] = (uint8_t)prev;
							}
						}
						if (bad || in.over || lens[256] == 0) { op = 3; break; }
						// lens[] is consumed by the first build before the distance lengths are moved down
						uint8_t dl[INF_MAXD];
						for (int k = 0; k < ndist; ++k) dl[k] = lens[nlen + k];
						int e = huff_build(hl, lens, nlen);
						if (e < 0 || (e > 0 && nlen - hl.count[0] != 1)) { op = 3; break; }
						for (int k = 0; k < ndist; ++k) lens[k] = dl[k];
						e = huff_build(hd, lens, ndist);
						if (e < 0 || (e > 0 && ndist - hd.count[0] != 1)) { op = 3; break; }
					} else { op = 3; break; }
					in_block = true;
				}
				// ---- symbols of a Huffman block
				const int sym = huff_decode(in, hl);
				if (sym < 0) { op = 3; break; }
				if (sym < 256) {
					if (out >= usize) { op = 3; break; }
					win[out++] = (uint8_t)sym;
				} else if (sym == 256) in_block = false;
				else {
					const int ls = sym - 257;
					if (ls >= 29) { op = 3; break; }
					const uint32_t len = c_lbase[ls] + in.bits(c_lext[ls]);
					const int ds = huff_decode(in, hd);
					if (ds < 0 || ds >= 30) { op = 3; break; }
					const uint32_t dist = c_dbase[ds] + in.bits(c_dext[ds]);
					if (in.over || dist > out || out + len > usize) { op = 3; break; }
					if (len < 8) { for (uint32_t i = 0; i < len; ++i) { win[out] = win[out - dist]; ++out; } continue; }   // not worth a hand-over
					s_a = (int)dist; s_b = (int)len; s_out = out;
					out += len;
					op = 1;
				}
			}
			s_op = op;
			status = op;
		}
		__syncwarp();
		const int op = s_op;
		if (op == 1) { // LZ77 copy by the warp: chunks of min(dist, 32) bytes never read what the same step writes
			const uint32_t dist = (uint32_t)s_a, len = (uint32_t)s_b, o = s_out;
			const uint32_t step = dist < 32u ? dist : 32u;
			for (uint32_t i = 0; i < len; i += step) {
				const uint32_t k = i + (uint32_t)lane;
				if ((uint32_t)lane < step && k < len) win[o + k] = win[o + k - dist];
				__syncwarp();
			}
		}
		__syncwarp();
		if (op >= 2) break;
	}
	status = s_op;
	const bool ok = status == 2 && __shfl_sync(0xffffffffu, out, 0) == usize;
	if (!ok && lane == 0) atomicOr(P.err, 256);
	// ---- window -> HBM (zeros if the block did not decode)
	uint8_t *dst = P.out + uoff;
	for (uint32_t i = lane; i < usize; i += 32) dst[i] = ok ? win[i] : 0;
}

cudaError_t launch_bgzf_inflate(const InflateParams &P, int n_blk, cudaStream_t st)
{
	if (n_blk <= 0) return cudaSuccess;
	const size_t smem = INF_WIN + ((size_t)P.max_csize + 31) / 16 * 16;
	cudaError_t e = cudaFuncSetAttribute(bgzf_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	bgzf_inflate_kernel<<<n_blk, 32, smem, st>>>(P);
	return cudaGetLastError();
}

} // namespace b200
