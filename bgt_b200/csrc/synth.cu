// synth.cu -- synthetic cohort generator: rows are drawn directly in PBWT-rank space, so the
// RLE stream of every row is known without any state; the running permutation is then obtained by running the
// rank-walk kernel in CHAIN mode over all rows and dumping S = rank^-1 in front of every checkpoint block, which
// makes the file truthful (the snapshots equal the encoder's state, pbwt.c:292-301).  The RLE is canonical
// (maximal runs, pbwt.c:24-50 digit split), so re-encoding the decoded matrix with the reference encoder
// reproduces the file byte for byte -- that is how tests validate this generator.
#include <cuda_runtime.h>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

struct Rng {
	uint64_t s;
	__device__ __forceinline__ explicit Rng(uint64_t seed) {
		uint64_t z = seed + 0x9E3779B97F4A7C15ull;          // splitmix64 to spread the (seed,row,plane) key
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		s = (z ^ (z >> 31)) | 1ull;
	}
	__device__ __forceinline__ uint64_t next() {           // xorshift64*
		s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
		return s * 0x2545F4914F6CDD1Dull;
	}
};

constexpr int SYNTH_MAX_IV = 64;

// intervals of 1s (start, length) of one (row, plane), unsorted
__device__ int draw_intervals(const SynthCfg &c, long long row, int plane, uint32_t *st, uint32_t *ln)
{
	Rng g(c.seed * 0xD1342543DE82EF95ull + (uint64_t)row * 2 + plane);
	const uint32_t m = c.m;
	int n;
	if (plane == 0) {
		int rmax = c.r_max < 1 ? 1 : (c.r_max > SYNTH_MAX_IV ? SYNTH_MAX_IV : c.r_max);
		n = 1 + (int)(g.next() % (uint64_t)rmax);
		const uint32_t cap = m / 2 ? m / 2 : 1;
		const int kmax = 31 - __clz(cap);
		// allele count of the row: uniform over octaves of [1, m/2] = log-uniform, i.e. a 1/x-like spectrum in
		// which rare variants dominate; it is spread over n intervals of 1s (fewer when the count is smaller)
		const int k = (int)(g.next() % (uint64_t)(kmax + 1));
		uint32_t ac = (1u << k) + (uint32_t)(g.next() % (1ull << k));
		if (ac > cap) ac = cap;
		if ((uint32_t)n > ac) n = (int)ac;
		const uint32_t span = 2 * (ac / (uint32_t)n) - 1;     // interval length uniform in [1, 2*ac/n - 1], mean ac/n
		for (int i = 0; i < n; ++i) {
			uint32_t len = 1 + (uint32_t)(g.next() % (uint64_t)span);
			if (len > cap) len = cap;
			ln[i] = len;
			st[i] = (uint32_t)(g.next() % (uint64_t)(m - len + 1));
		}
	} else {
		const int one_in = c.p1_one_in < 1 ? 1 : c.p1_one_in;
		if (g.next() % (uint64_t)one_in) return 0;
		const int max_iv = c.p1_max_iv < 1 ? 3 : (c.p1_max_iv > SYNTH_MAX_IV ? SYNTH_MAX_IV : c.p1_max_iv);
		const uint32_t max_len = c.p1_max_len < 1 ? 64u : (uint32_t)c.p1_max_len;
		n = 1 + (int)(g.next() % (uint64_t)max_iv);
		for (int i = 0; i < n; ++i) {
			uint32_t len = 1 + (uint32_t)(g.next() % (uint64_t)max_len);
			if (len > m) len = m;
			ln[i] = len;
			st[i] = (uint32_t)(g.next() % (uint64_t)(m - len + 1));
		}
	}
	return n;
}

// pbwt.c:24-36
__device__ __forceinline__ uint32_t put_run(uint8_t *p, uint32_t len, uint32_t bit)
{
	if (len < 16) { if (p) p[0] = (uint8_t)(len << 1 | bit); return 1; }
	uint32_t n = 0;
	for (int pos = 7; pos >= 0; --pos) {
		const uint32_t d = (len >> (4 * pos)) & 15u;
		if (d) { if (p) p[n] = (uint8_t)((((uint32_t)pos << 4) | d) << 1 | bit); ++n; }
	}
	return n;
}

// canonical RLE of one (row, plane); p == nullptr only counts
__device__ uint32_t emit_plane(const SynthCfg &c, long long row, int plane, uint8_t *p)
{
	uint32_t st[SYNTH_MAX_IV], ln[SYNTH_MAX_IV];
	const int n = draw_intervals(c, row, plane, st, ln);
	for (int i = 1; i < n; ++i) { // insertion sort by start
		const uint32_t s = st[i], l = ln[i];
		int j = i - 1;
		while (j >= 0 && st[j] > s) { st[j + 1] = st[j]; ln[j + 1] = ln[j]; --j; }
		st[j + 1] = s; ln[j + 1] = l;
	}
	uint32_t pos = 0, out = 0;
	int i = 0;
	while (i < n) {
		uint32_t s = st[i], e = st[i] + ln[i];
		++i;
		while (i < n && st[i] <= e) { const uint32_t e2 = st[i] + ln[i]; if (e2 > e) e = e2; ++i; } // merge overlapping / adjacent
		if (s > pos) out += put_run(p ? p + out : nullptr, s - pos, 0);
		out += put_run(p ? p + out : nullptr, e - s, 1);
		pos = e;
	}
	if (pos < c.m) out += put_run(p ? p + out : nullptr, c.m - pos, 0);
	return out;
}

__global__ void __launch_bounds__(128) synth_lengths_kernel(const SynthCfg c, uint32_t *__restrict__ len2)
{
	const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= c.n_rows) return;
	len2[row * 2 + 0] = emit_plane(c, row, 0, nullptr);
	len2[row * 2 + 1] = emit_plane(c, row, 1, nullptr);
}

__global__ void __launch_bounds__(128) synth_write_kernel(const SynthCfg c, const uint64_t *__restrict__ rowoff, uint8_t *__restrict__ img)
{
	const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= c.n_rows) return;
	uint8_t *p = img + rowoff[row];
	*p++ = 'B';
	for (int plane = 0; plane < 2; ++plane) {
		const uint32_t l = emit_plane(c, row, plane, p + 4);
		p[0] = (uint8_t)l; p[1] = (uint8_t)(l >> 8); p[2] = (uint8_t)(l >> 16); p[3] = (uint8_t)(l >> 24);
		p += 4 + l;
	}
}

cudaError_t launch_synth_lengths(const SynthCfg &c, uint32_t *len2, cudaStream_t st)
{
	if (c.n_rows <= 0) return cudaSuccess;
	synth_lengths_kernel<<<(unsigned)((c.n_rows + 127) / 128), 128, 0, st>>>(c, len2);
	return cudaGetLastError();
}

cudaError_t launch_synth_write(const SynthCfg &c, const uint64_t *rowoff_flat, uint8_t *img, cudaStream_t st)
{
	if (c.n_rows <= 0) return cudaSuccess;
	synth_write_kernel<<<(unsigned)((c.n_rows + 127) / 128), 128, 0, st>>>(c, rowoff_flat, img);
	return cudaGetLastError();
}

} // namespace b200
