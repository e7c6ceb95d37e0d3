/*
 * pbc_host.c -- the in-memory codec entry points of pbwt.h (pbwt.h:98-130) for seam A, so that a host application can
 * drop the reference's pbwt.o entirely: pbc_init, pbc_enc, pbc_dec, pbs_dec and the two undeclared-but-visible cores
 * pbc_enc_core / pbc_dec_core (pbwt.c:57,69).
 *
 * Nothing on the genotype hot path calls these (the file API pbf_* is served by the GPU, pbwt_shim.c): in the reference
 * they have no caller outside pbwt.c.  They operate on ONE row held in host memory through the caller-visible pbc_t
 * {m, l, S0, S, u} -- a single row of a few hundred KB is not worth a PCIe round trip -- so they are plain C with the
 * reference's exact semantics (same S, same u bytes, same NUL terminator), written from the definitions in
 * SURVEY App. A/B rather than from the reference's loops.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/pbwt_b200.h"

/* run length of code byte c (pbwt.c:12-21 as arithmetic: v = c>>1, low nibble = hex digit, high bits = its position) */
static inline uint32_t code_len(uint8_t c)
{
	const uint32_t v = c >> 1;
	return (v & 15u) << ((v >> 4) << 2);
}

/* a run of l equal bits -> one byte per non-zero hex digit of l, most significant first (pbwt.c:24-36); returns bytes written */
static inline int put_run(uint8_t *dst, uint32_t l, uint32_t bit)
{
	int k = 0, pos;
	for (pos = 7; pos >= 0; --pos) {
		const uint32_t digit = (l >> (4 * pos)) & 15u;
		if (digit) dst[k++] = (uint8_t)(((uint32_t)pos << 4 | digit) << 1 | bit);
	}
	return k;
}

static uint32_t ones_of(const uint8_t *u)
{
	uint32_t n1 = 0;
	for (; *u; ++u) if (*u & 1) n1 += code_len(*u);
	return n1;
}

/* A_k in column order + S_{k-1} -> B_k as RLE in u (NUL terminated, u must hold m+1 bytes) + S_k; returns strlen(u) (pbwt.c:57-66) */
int pbc_enc_core(int m, const int32_t *S0, const uint8_t *a, int32_t *S, uint8_t *u)
{
	int j, n1 = 0, z = 0, o, w = 0;
	uint32_t run = 0, cur = 0;
	for (j = 0; j < m; ++j) n1 += (a[S0[j]] != 0);
	o = m - n1;
	for (j = 0; j < m; ++j) { /* rank order: bit of the j-th column of S_{k-1}; stable partition; runs coded on the fly */
		const int32_t col = S0[j];
		const uint32_t bit = a[col] != 0;
		if (bit) S[o++] = col; else S[z++] = col;
		if (run && bit != cur) { w += put_run(u + w, run, cur); run = 0; }
		cur = bit; ++run;
	}
	if (run) w += put_run(u + w, run, cur);
	u[w] = 0;
	return w;
}

/* B_k (RLE, NUL terminated) + S_{k-1} -> A_k in column order + S_k (pbwt.c:69-90) */
void pbc_dec_core(int m, const int32_t *S0, const uint8_t *u, int32_t *S, uint8_t *a)
{
	const uint32_t n1 = ones_of(u);
	int z = 0, o = m - (int)n1, s = 0;
	if (n1 == 0 || n1 == (uint32_t)m) { /* constant row: the order does not change */
		memcpy(S, S0, (size_t)m * sizeof(int32_t));
		memset(a, n1 != 0, (size_t)m);
		return;
	}
	memset(a, 0, (size_t)m);
	for (; *u; ++u) {
		const int e = s + (int)code_len(*u);
		if (*u & 1) for (; s < e; ++s) { S[o++] = S0[s]; a[S0[s]] = 1; }
		else        for (; s < e; ++s) S[z++] = S0[s];
	}
}

pbc_t *pbc_init(int m) /* one allocation, released with free() (pbwt.c:93-105) */
{
	uint8_t *p = (uint8_t*)calloc(sizeof(pbc_t) + 8 * (size_t)m + (size_t)m + 1, 1);
	pbc_t *pb = (pbc_t*)p;
	int j;
	if (p == 0) return 0;
	pb->m = m;
	pb->S0 = (int32_t*)(p + sizeof(pbc_t));
	pb->S = pb->S0 + m;
	pb->u = (uint8_t*)(pb->S + m);
	for (j = 0; j < m; ++j) pb->S[j] = j;
	return pb;
}

void pbc_enc(pbc_t *pb, const uint8_t *a) /* pbwt.c:107-112 */
{
	int32_t *t = pb->S; pb->S = pb->S0; pb->S0 = t;
	pb->l = pbc_enc_core(pb->m, pb->S0, a, pb->S, pb->u);
}

void pbc_dec(pbc_t *pb, const uint8_t *b) /* pbwt.c:114-119 */
{
	int32_t *t = pb->S; pb->S = pb->S0; pb->S0 = t;
	pbc_dec_core(pb->m, pb->S0, b, pb->S, pb->u);
}

/* Subset decode (pbwt.c:129-170): d[] = (rank, output slot) sorted by rank.  Every entry moves to
 * rank' = rank + delta(run) and a[slot] = the run's bit; afterwards d is sorted by rank' again, which a
 * stable split into "landed among the zeros" / "landed among the ones" gives for free. */
void pbs_dec(int m, int r, pbs_dat_t *d, const uint8_t *u, uint8_t *a)
{
	const uint32_t n1 = ones_of(u);
	uint32_t start = 0, before[2] = {0, 0}, base[2];
	pbs_dat_t *ones;
	int i = 0, nz = 0, no = 0;
	if (r <= 0) return;
	if (n1 == 0 || n1 == (uint32_t)m) { memset(a, n1 != 0, (size_t)r); return; }
	base[0] = 0; base[1] = (uint32_t)m - n1;
	ones = (pbs_dat_t*)malloc((size_t)r * sizeof(pbs_dat_t));
	memset(a, 0, (size_t)r);
	for (; *u && i < r; ++u) {
		const uint32_t len = code_len(*u), bit = *u & 1u, end = start + len;
		const uint32_t shift = base[bit] + before[bit] - start;   /* rank' - rank inside this run (mod 2^32) */
		for (; i < r && d[i].r < end; ++i) {
			pbs_dat_t e = d[i];
			e.r += shift;
			if (bit) { ones[no++] = e; a[e.i] = 1; }
			else d[nz++] = e;                                      /* nz <= i: in place */
		}
		before[bit] += len;
		start = end;
	}
	memcpy(d + nz, ones, (size_t)no * sizeof(pbs_dat_t));
	free(ones);
}
