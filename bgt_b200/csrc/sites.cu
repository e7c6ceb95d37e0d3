// sites.cu -- the site side of `bgt view` on the device: BCF records -> site table -> VCF text.
//
// What the reference does per site on the host thread: bcf_read1 + bcf_unpack of the site-only record
// (bgt.c:272-288, vcf.c:316-360), bcfcpy_min (vcf.c:1166-1182), bgtm_fill_info (bgt.c:721-733), vcf_format1
// (vcf.c:895-969).  Here the inflated .bcf stream (inflate.cu) is indexed and parsed by kernels:
//   bcf_chase_kernel   record offsets: records are length-prefixed; the record-number index of the .csi (RNI, one BGZF
//                      virtual offset per 1024 records, hts.c:394-400,536-542) gives independent starting points, one
//                      thread chases each stretch.
//   bcf_parse_kernel   one thread per record: rid, pos, rlen, n_allele, where REF and the first ALT lie, INFO/_row.
//   view_len/_write    the VCF line of every site that passes the filter, from the scan's per-row counts: lengths,
//                      exclusive prefix sum (CUB), then the bytes -- identical to vcf_format1's for such a record.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <stdint.h>
#include "pbwt_kernels.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t rd_u32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

// ------------------------------------------------------------------------------------------------ record offsets

// One thread per stretch: seg_pos[k] = offset of record k*seg_len in the stream.  count_only: the single-stretch fallback
// without an RNI first counts the records up to the end of the stream.
__global__ void __launch_bounds__(128) bcf_chase_kernel(const uint8_t *__restrict__ bcf, unsigned long long bcf_len, const unsigned long long *__restrict__ seg_pos,
                                                        int n_seg, int seg_len, long long n_rec, unsigned long long *__restrict__ rec_off,
                                                        unsigned long long *__restrict__ counted, int *__restrict__ err)
{
	const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	if (k >= n_seg) return;
	unsigned long long pos = seg_pos[k];
	if (counted) { // count records from pos to the end of the stream
		unsigned long long n = 0;
		while (pos + 32 <= bcf_len) {
			const unsigned long long nx = pos + 8ull + rd_u32(bcf + pos) + rd_u32(bcf + pos + 4);
			if (nx > bcf_len || rd_u32(bcf + pos) < 24u) { atomicOr(err, 512); break; }
			pos = nx; ++n;
		}
		if (pos != bcf_len) atomicOr(err, 512);
		*counted = n;
		return;
	}
	const long long r0 = (long long)k * seg_len;
	const long long r1 = r0 + seg_len < n_rec ? r0 + seg_len : n_rec;
	for (long long r = r0; r < r1; ++r) {
		if (pos + 32 > bcf_len || rd_u32(bcf + pos) < 24u) { atomicOr(err, 512); for (; r < r1; ++r) rec_off[r] = ~0ull; return; }
		rec_off[r] = pos;
		pos += 8ull + rd_u32(bcf + pos) + rd_u32(bcf + pos + 4);
	}
	if (pos > bcf_len) atomicOr(err, 512);
}

cudaError_t launch_bcf_chase(const uint8_t *bcf, unsigned long long bcf_len, const unsigned long long *seg_pos, int n_seg, int seg_len, long long n_rec,
                             unsigned long long *rec_off, unsigned long long *counted, int *err, cudaStream_t st)
{
	if (n_seg <= 0) return cudaSuccess;
	bcf_chase_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(bcf, bcf_len, seg_pos, n_seg, seg_len, n_rec, rec_off, counted, err);
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ record parse

// typed-value descriptor (vcf.c:430-459 writers; BCF2 spec): low nibble type (1 int8, 2 int16, 3 int32, 5 float, 7 char),
// high nibble length, 15 = the length follows as a typed integer
__device__ __forceinline__ bool bcf_desc(const uint8_t *&q, const uint8_t *end, int &type, int &len)
{
	if (q >= end) return false;
	const uint8_t d = *q++;
	type = d & 15; len = d >> 4;
	if (len == 15) {
		if (q >= end) return false;
		const int t2 = *q & 15;
		++q;
		if (t2 == 1) { if (q + 1 > end) return false; len = (int8_t)q[0]; q += 1; }
		else if (t2 == 2) { if (q + 2 > end) return false; len = (int16_t)(q[0] | q[1] << 8); q += 2; }
		else if (t2 == 3) { if (q + 4 > end) return false; len = (int32_t)rd_u32(q); q += 4; }
		else return false;
		if (len < 0) return false;
	}
	return true;
}
__device__ __forceinline__ int bcf_tsize(int type) { return type == 1 || type == 7 ? 1 : type == 2 ? 2 : (type == 3 || type == 5) ? 4 : 0; }
__device__ __forceinline__ long long bcf_int(const uint8_t *q, int type)
{
	return type == 1 ? (long long)(int8_t)q[0] : type == 2 ? (long long)(int16_t)(q[0] | q[1] << 8) : (long long)(int32_t)rd_u32(q);
}

__global__ void __launch_bounds__(256) bcf_parse_kernel(const uint8_t *__restrict__ bcf, unsigned long long bcf_len, const unsigned long long *__restrict__ rec_off,
                                                        long long n_rec, int row_key, SiteRec *__restrict__ sites, int *__restrict__ err)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rec) return;
	SiteRec s;
	s.rid = s.pos = s.rlen = 0; s.n_allele = 0; s.ref_off = s.alt_off = 0; s.ref_len = s.alt_len = 0; s.row = -1;
	const unsigned long long off = rec_off[i];
	bool ok = off != ~0ull && off + 32 <= bcf_len;
	if (ok) {
		const uint8_t *p = bcf + off;
		const uint32_t l_shared = rd_u32(p);
		const uint8_t *end = p + 8 + l_shared;
		ok = l_shared >= 24 && off + 8ull + l_shared <= bcf_len;
		if (ok) {
			s.rid = (int32_t)rd_u32(p + 8); s.pos = (int32_t)rd_u32(p + 12); s.rlen = (int32_t)rd_u32(p + 16);
			const uint32_t w = rd_u32(p + 24);
			const int n_info = (int)(w & 0xffffu);
			s.n_allele = (int)(w >> 16);
			const uint8_t *q = p + 32;
			int type, len;
			ok = bcf_desc(q, end, type, len) && type == 7 && q + len <= end;          // ID
			if (ok) q += len;
			for (int a = 0; ok && a < s.n_allele; ++a) {                              // alleles
				ok = bcf_desc(q, end, type, len) && (type == 7 || len == 0) && q + len <= end;
				if (!ok) break;
				if (a == 0) { s.ref_off = (unsigned long long)(q - bcf); s.ref_len = len; }
				else if (a == 1) { s.alt_off = (unsigned long long)(q - bcf); s.alt_len = len; }
				q += len;
			}
			if (ok) { ok = bcf_desc(q, end, type, len) && q + (size_t)len * bcf_tsize(type) <= end; if (ok) q += (size_t)len * bcf_tsize(type); }   // FILTER
			for (int k = 0; ok && k < n_info; ++k) {                                 // INFO: key, value
				ok = bcf_desc(q, end, type, len) && len == 1 && type >= 1 && type <= 3 && q + bcf_tsize(type) <= end;
				if (!ok) break;
				const long long key = bcf_int(q, type);
				q += bcf_tsize(type);
				ok = bcf_desc(q, end, type, len) && q + (size_t)len * bcf_tsize(type) <= end;
				if (!ok) break;
				if ((row_key < 0 ? k == 0 : key == row_key) && len >= 1 && type >= 1 && type <= 3) s.row = bcf_int(q, type);   // bgt.c:279-286
				q += (size_t)len * bcf_tsize(type);
			}
			ok = ok && s.row >= 0 && s.n_allele >= 2;                                // bgt.c:277,280,286 assert these
		}
	}
	if (!ok) { atomicOr(err, 1024); s.row = -1; }
	sites[i] = s;
}

cudaError_t launch_bcf_parse(const uint8_t *bcf, unsigned long long bcf_len, const unsigned long long *rec_off, long long n_rec, int row_key, SiteRec *sites,
                             int *err, cudaStream_t st)
{
	if (n_rec <= 0) return cudaSuccess;
	bcf_parse_kernel<<<(unsigned)((n_rec + 255) / 256), 256, 0, st>>>(bcf, bcf_len, rec_off, n_rec, row_key, sites, err);
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ VCF text

__device__ __forceinline__ int dec_len(long long v) // characters kputw (kstring.h:112-127) prints
{
	int n = v < 0 ? 1 : 0;
	unsigned long long x = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
	do { ++n; x /= 10; } while (x);
	return n;
}
__device__ __forceinline__ char *put_dec(char *q, long long v)
{
	char tmp[24];
	int n = 0;
	unsigned long long x = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
	do { tmp[n++] = (char)('0' + x % 10); x /= 10; } while (x);
	if (v < 0) *q++ = '-';
	while (n) *q++ = tmp[--n];
	return q;
}
__device__ __forceinline__ char *put_str(char *q, const char *s, int n) { for (int i = 0; i < n; ++i) q[i] = s[i]; return q + n; }

// The line of one site (without GT columns: `view -G`), or its length when out == nullptr.
//   CHROM \t POS \t . \t REF \t ALT[,<M>] \t 0 \t . \t [END=e;]AN=..;AC=..[,..][;AN1=..;AC1=..[,..] ...] \n
// (bcfcpy_min: empty ID, first ALT, <M> iff the stored record has more than two alleles, QUAL 0, no FILTER;
//  bgt.c:824-827 END; bgtm_fill_info: AC has n_allele-1 values, group keys only with several groups.)
__device__ long long site_line(const ViewParams &P, const SiteRec &s, const int32_t *c, char *out, bool nl)
{
	const bool multi = s.n_allele > 2;
	const bool has_end = s.ref_len != s.rlen;
	const int rid = s.rid >= 0 && s.rid < P.n_ctg ? s.rid : 0;
	const int ctg_len = P.ctg_off[rid + 1] - P.ctg_off[rid];
	if (!out) {
		long long n = ctg_len + 1 + dec_len((long long)s.pos + 1) + 3 + s.ref_len + 1 + s.alt_len + (multi ? 4 : 0) + 5;   // ... \t0\t.\t
		if (has_end) n += 4 + dec_len((long long)s.pos + s.rlen) + (P.with_counts ? 1 : 0);
		if (P.with_counts) {
			n += 3 + dec_len(c[0]) + 4 + dec_len(c[1]) + (multi ? 1 + dec_len(c[2]) : 0);
			if (P.G > 1)
				for (int g = 0; g < P.G; ++g) {
					const int kl = g < 9 ? 3 : 4;
					n += 1 + kl + 1 + dec_len(c[3 + 3 * g]) + 1 + kl + 1 + dec_len(c[4 + 3 * g]) + (multi ? 1 + dec_len(c[5 + 3 * g]) : 0);
				}
		} else if (!has_end) n += 1;   // "."
		return n + (nl ? 1 : 0);
	}
	char *q = out;
	q = put_str(q, P.ctg_names + P.ctg_off[rid], ctg_len);
	*q++ = '\t'; q = put_dec(q, (long long)s.pos + 1);
	*q++ = '\t'; *q++ = '.'; *q++ = '\t';
	q = put_str(q, (const char*)P.bcf + s.ref_off, s.ref_len);
	*q++ = '\t';
	q = put_str(q, (const char*)P.bcf + s.alt_off, s.alt_len);
	if (multi) { *q++ = ','; *q++ = '<'; *q++ = 'M'; *q++ = '>'; }
	*q++ = '\t'; *q++ = '0'; *q++ = '\t'; *q++ = '.'; *q++ = '\t';
	if (has_end) { *q++ = 'E'; *q++ = 'N'; *q++ = 'D'; *q++ = '='; q = put_dec(q, (long long)s.pos + s.rlen); if (P.with_counts) *q++ = ';'; }
	if (P.with_counts) {
		*q++ = 'A'; *q++ = 'N'; *q++ = '='; q = put_dec(q, c[0]);
		*q++ = ';'; *q++ = 'A'; *q++ = 'C'; *q++ = '='; q = put_dec(q, c[1]);
		if (multi) { *q++ = ','; q = put_dec(q, c[2]); }
		if (P.G > 1)
			for (int g = 0; g < P.G; ++g) {
				for (int w = 0; w < 2; ++w) { // gen_group_key (bgt.c:692-698): AN<g+1>, AC<g+1>
					*q++ = ';'; *q++ = 'A'; *q++ = w ? 'C' : 'N';
					if (g < 9) *q++ = (char)('0' + g + 1); else { *q++ = (char)('0' + (g + 1) / 10); *q++ = (char)('0' + (g + 1) % 10); }
					*q++ = '=';
					q = put_dec(q, c[3 + 3 * g + w]);
					if (w && multi) { *q++ = ','; q = put_dec(q, c[5 + 3 * g]); }
				}
			}
	} else if (!has_end) *q++ = '.';
	if (nl) *q++ = '\n';
	return (long long)(q - out);
}

__global__ void __launch_bounds__(256) view_len_kernel(const ViewParams P, unsigned long long *__restrict__ len)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_rec) return;
	const SiteRec s = P.sites[i];
	const long long r = s.row - P.row_lo;
	unsigned long long n = 0;
	if (s.row >= 0 && r >= 0 && r < P.n_rows) {
		if (!P.pass || P.pass[r]) {
			n = (unsigned long long)site_line(P, s, P.counts ? P.counts + (size_t)r * P.stride : nullptr, nullptr, true);
			if (P.with_gt) n += 3ull + 4ull * (unsigned long long)P.n_out;   // "\tGT" + "\ta/b" per sample (bgt_gen_gt bgt.c:290-313, vcf.c:940-966)
		}
	} else if (s.row >= 0) atomicOr(P.err, 2048);   // a site points at a row that was not scanned
	len[i] = n;
}

__global__ void __launch_bounds__(256) view_write_kernel(const ViewParams P, const unsigned long long *__restrict__ len, const unsigned long long *__restrict__ off,
                                                         char *__restrict__ text, unsigned long long *__restrict__ n_lines)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long wrote = 0;
	if (i < P.n_rec && len[i]) {
		const SiteRec s = P.sites[i];
		const long long r = s.row - P.row_lo;
		site_line(P, s, P.counts ? P.counts + (size_t)r * P.stride : nullptr, text + off[i], true);
		wrote = 1;
	}
	wrote = __reduce_add_sync(0xffffffffu, (unsigned)wrote);
	if ((threadIdx.x & 31) == 0 && wrote) atomicAdd(n_lines, wrote);
}

// With genotype columns: one CTA per record.  Thread 0 writes the fixed columns and "\tGT"; all threads write the samples'
// "\ta/b" from the scan's bit planes: code a1<<1|a0 -> bgt_bits2gt = {0, 1, ., 2} (bgt.c:250), always unphased.
__global__ void __launch_bounds__(128) view_write_gt_kernel(const ViewParams P, const unsigned long long *__restrict__ len, const unsigned long long *__restrict__ off,
                                                            char *__restrict__ text, unsigned long long *__restrict__ n_lines)
{
	const long long i = blockIdx.x;
	const unsigned long long L = len[i];
	if (L == 0) return;
	const SiteRec s = P.sites[i];
	const long long r = s.row - P.row_lo;
	char *line = text + off[i];
	const unsigned long long gt_bytes = 3ull + 4ull * (unsigned long long)P.n_out;
	char *g = line + (L - 1 - gt_bytes);
	if (threadIdx.x == 0) {
		site_line(P, s, P.counts ? P.counts + (size_t)r * P.stride : nullptr, line, false);
		g[0] = '\t'; g[1] = 'G'; g[2] = 'T';
		line[L - 1] = '\n';
		atomicAdd(n_lines, 1ull);
	}
	const uint32_t *h0 = P.hap[0] + (size_t)r * P.words, *h1 = P.hap[1] + (size_t)r * P.words;
	g += 3;
	for (int smp = threadIdx.x; smp < P.n_out; smp += 128) {
		const int c0 = 2 * smp, w = c0 >> 5, b = c0 & 31;          // the sample's two haplotypes share a word (c0 is even)
		const uint32_t v0 = h0[w] >> b, v1 = h1[w] >> b;
		const uint32_t ca = (v0 & 1u) | (v1 & 1u) << 1, cb = ((v0 >> 1) & 1u) | ((v1 >> 1) & 1u) << 1;
		char *q = g + 4 * (size_t)smp;
		q[0] = '\t'; q[1] = "01.2"[ca]; q[2] = '/'; q[3] = "01.2"[cb];
	}
}

size_t view_scan_temp_bytes(long long n)
{
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)(n + 1));
	return bytes;
}

cudaError_t launch_view_text(const ViewParams &P, unsigned long long *len, unsigned long long *off, void *temp, size_t temp_bytes, char *text,
                             unsigned long long *n_lines, int phase, cudaStream_t st)
{
	if (P.n_rec <= 0) return cudaSuccess;
	const unsigned grid = (unsigned)((P.n_rec + 255) / 256);
	if (phase == 0) { // lengths + offsets (entry n_rec of `off` = total: len has one trailing zero entry)
		view_len_kernel<<<grid, 256, 0, st>>>(P, len);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return e;
		return cub::DeviceScan::ExclusiveSum(temp, temp_bytes, len, off, (int)(P.n_rec + 1), st);
	}
	if (P.with_gt) view_write_gt_kernel<<<(unsigned)P.n_rec, 128, 0, st>>>(P, len, off, text, n_lines);
	else view_write_kernel<<<grid, 256, 0, st>>>(P, len, off, text, n_lines);
	return cudaGetLastError();
}

} // namespace b200
