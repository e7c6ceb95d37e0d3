/*
 * pbwt_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * CPU restatement of the reference PBWT codec and PBF container (pbwt.c), memory-resident.
 * Pinned against oracle/_ref (the unmodified reference) by tests/test_oracle_pin.py.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "oracle.h"

/* ------------------------------------------------------------------ RLE alphabet */

/* pbwt.c:12-21: entry v = c>>1 of the 128-entry table is (v&15) << 4*(v>>4). */
uint32_t orc_rle_len(uint8_t c)
{
	uint32_t v = c >> 1;
	return (v & 15u) << (4 * (v >> 4));
}

/* pbwt.c:24-36: a run < 16 is one byte; otherwise one byte per non-zero hex digit of the length,
 * most significant digit first; byte = (digit_position<<4 | digit) << 1 | bit. */
int orc_rle_put_run(uint8_t *p, uint32_t len, int bit)
{
	int n = 0, pos;
	if (len < 16) { p[0] = (uint8_t)(len << 1 | bit); return 1; }
	for (pos = 7; pos >= 0; --pos) {
		uint32_t d = (len >> (4 * pos)) & 15u;
		if (d) p[n++] = (uint8_t)(((uint32_t)pos << 4 | d) << 1 | bit);
	}
	return n;
}

/* pbwt.c:39-50: maximal runs of equal bytes, NUL-terminated. out may alias bits (as in the reference). */
int orc_rle_encode(int m, const uint8_t *bits, uint8_t *out)
{
	uint8_t *p = out, last = bits[0];
	int j, l = 1;
	for (j = 1; j < m; ++j) {
		uint8_t x = bits[j];
		if (x == last) ++l;
		else { p += orc_rle_put_run(p, l, last); l = 1; last = x; }
	}
	p += orc_rle_put_run(p, l, last);
	*p = 0;
	return (int)(p - out);
}

/* ------------------------------------------------------------------ full codec */

orc_codec_t *orc_codec_new(int m)
{
	orc_codec_t *c = (orc_codec_t*)calloc(1, sizeof(*c));
	int j;
	c->m = m;
	c->cur = (int32_t*)calloc(m > 0 ? m : 1, 4);
	c->prev = (int32_t*)calloc(m > 0 ? m : 1, 4);
	c->u = (uint8_t*)calloc((size_t)m + 1, 1);
	for (j = 0; j < m; ++j) c->cur[j] = j; /* pbwt.c:103 identity before row 0 */
	return c;
}

void orc_codec_free(orc_codec_t *c)
{
	if (!c) return;
	free(c->cur); free(c->prev); free(c->u); free(c);
}

void orc_codec_set_perm(orc_codec_t *c, const int32_t *S) { memcpy(c->cur, S, (size_t)c->m * 4); }

/* pbwt.c:57-66 + 107-112: B_k[j] = !!A_k[S_{k-1}[j]]; S_k = stable partition of S_{k-1} by B_k. */
void orc_codec_enc(orc_codec_t *c, const uint8_t *a)
{
	int32_t *t = c->cur, *z, *o;
	int j, n1 = 0, m = c->m;
	c->cur = c->prev; c->prev = t;
	for (j = 0; j < m; ++j) n1 += (c->u[j] = (a[c->prev[j]] != 0));
	z = c->cur; o = c->cur + (m - n1);
	for (j = 0; j < m; ++j) {
		if (c->u[j]) *o++ = c->prev[j];
		else *z++ = c->prev[j];
	}
	c->l = orc_rle_encode(m, c->u, c->u);
}

/* pbwt.c:69-90 + 114-119 */
void orc_codec_dec(orc_codec_t *c, const uint8_t *rle)
{
	int32_t *t = c->cur, *dst[2];
	const uint8_t *q;
	int64_t n1 = 0;
	int m = c->m, s;
	c->cur = c->prev; c->prev = t;
	for (q = rle; *q; ++q) if (*q & 1) n1 += orc_rle_len(*q);
	if (n1 == 0 || n1 == m) { /* pbwt.c:75-77: identity permutation, constant row */
		memcpy(c->cur, c->prev, (size_t)m * 4);
		memset(c->u, n1 == m, m);
		return;
	}
	dst[0] = c->cur; dst[1] = c->cur + (m - n1);
	memset(c->u, 0, m);
	for (q = rle, s = 0; *q; ++q) {
		uint32_t len = orc_rle_len(*q), i;
		int b = *q & 1;
		const int32_t *src = c->prev + s;
		for (i = 0; i < len; ++i) {
			if (b) c->u[src[i]] = 1;
			*dst[b]++ = src[i];
		}
		s += len;
	}
}

/* ------------------------------------------------------------------ subset codec */

static int sub_cmp(const void *x, const void *y)
{
	uint32_t a = ((const orc_sub_t*)x)->r, b = ((const orc_sub_t*)y)->r;
	return (a > b) - (a < b);
}

/* pbwt.c:340-347: rank of each wanted column under S, then sort by rank (ranks are distinct, so any
 * sort reproduces the reference's radix sort). d[i].i must already hold the output slot. */
void orc_subset_fill(int m, const int32_t *S, int n_sub, orc_sub_t *d, const int *cols)
{
	int32_t *inv = (int32_t*)malloc((size_t)m * 4);
	int i;
	for (i = 0; i < m; ++i) inv[S[i]] = i;
	for (i = 0; i < n_sub; ++i) d[i].r = inv[cols[d[i].i]];
	qsort(d, n_sub, sizeof(orc_sub_t), sub_cmp);
	free(inv);
}

/* pbwt.c:129-170: merge-walk of the runs and the rank-sorted entries; an entry in run (start s, bit b)
 * moves to acc[b] + c[b] + (r - s); zeros stay in front, ones are appended, so d stays rank-sorted. */
void orc_subset_dec(int m, int n_sub, orc_sub_t *d, const uint8_t *rle, uint8_t *a)
{
	const uint8_t *q;
	int64_t n1 = 0;
	for (q = rle; *q; ++q) if (*q & 1) n1 += orc_rle_len(*q);
	if (n1 == 0 || n1 == m) { memset(a, n1 == m && n1 != 0, n_sub); return; }
	{
		orc_sub_t *ones = (orc_sub_t*)malloc((size_t)(n_sub ? n_sub : 1) * sizeof(orc_sub_t));
		int64_t cnt[2] = {0, 0}, acc[2];
		int p = 0, nz = 0, no = 0;
		acc[0] = 0; acc[1] = m - n1;
		memset(a, 0, n_sub);
		for (q = rle; *q && p < n_sub; ++q) {
			int64_t len = orc_rle_len(*q), s = cnt[0] + cnt[1];
			int b = *q & 1;
			while (p < n_sub && (int64_t)d[p].r >= s && (int64_t)d[p].r < s + len) {
				orc_sub_t e = d[p++];
				e.r = (uint32_t)(acc[b] + cnt[b] + ((int64_t)e.r - s));
				if (b) { ones[no++] = e; a[e.i] = 1; }
				else d[nz++] = e; /* nz <= p-1, in-place compaction is safe */
			}
			cnt[b] += len;
		}
		/* entries past the last run (cannot happen on a well-formed row) keep the reference's layout:
		 * pbwt.c:167 appends the ones right after the zeros written so far */
		memcpy(d + nz, ones, (size_t)no * sizeof(orc_sub_t));
		free(ones);
	}
}

/* ------------------------------------------------------------------ PBF reader */

struct orc_pbf_s {
	const uint8_t *buf;
	size_t len, pos;
	int32_t m, g, shift;
	int64_t n, k;          /* rows in file; next row to read */
	int32_t n_idx;
	uint64_t *idx;
	orc_codec_t **pc;
	const uint8_t **ret;
	int n_sub;
	int *cols;
	orc_sub_t **sub;
	uint8_t *tmp;
};

static int rd(orc_pbf_t *p, void *dst, size_t n)
{
	if (p->pos + n > p->len) { memset(dst, 0, n); p->pos = p->len; return -1; }
	memcpy(dst, p->buf + p->pos, n); p->pos += n;
	return 0;
}

/* pbwt.c:221-262: "PBF\1", m, g, shift; trailing 8 bytes point at the 'I' record */
orc_pbf_t *orc_pbf_open_mem(const uint8_t *buf, size_t len)
{
	orc_pbf_t *p;
	int32_t v[3];
	int i;
	if (len < 16 || memcmp(buf, "PBF\1", 4) != 0) return 0;
	p = (orc_pbf_t*)calloc(1, sizeof(*p));
	p->buf = buf; p->len = len;
	memcpy(v, buf + 4, 12);
	p->m = v[0]; p->g = v[1]; p->shift = v[2];
	p->pc = (orc_codec_t**)calloc(p->g, sizeof(void*));
	p->ret = (const uint8_t**)calloc(p->g, sizeof(void*));
	p->sub = (orc_sub_t**)calloc(p->g, sizeof(void*));
	for (i = 0; i < p->g; ++i) { p->pc[i] = orc_codec_new(p->m); p->ret[i] = p->pc[i]->u; }
	p->tmp = (uint8_t*)calloc((size_t)p->m + 1, 1);
	if (len >= 24) {
		uint64_t off;
		memcpy(&off, buf + len - 8, 8);
		if (off + 13 <= len && buf[off] == 'I') {
			memcpy(&p->n, buf + off + 1, 8);
			memcpy(&p->n_idx, buf + off + 9, 4);
			p->idx = (uint64_t*)calloc(p->n_idx > 0 ? p->n_idx : 1, 8);
			memcpy(p->idx, buf + off + 13, (size_t)p->n_idx * 8);
		}
	}
	p->pos = 16;
	return p;
}

void orc_pbf_close(orc_pbf_t *p)
{
	int i;
	if (!p) return;
	for (i = 0; i < p->g; ++i) { orc_codec_free(p->pc[i]); free(p->sub[i]); }
	free(p->pc); free(p->ret); free(p->sub); free(p->idx); free(p->cols); free(p->tmp); free(p);
}

int orc_pbf_m(const orc_pbf_t *p) { return p->m; }
int orc_pbf_g(const orc_pbf_t *p) { return p->g; }
int orc_pbf_shift(const orc_pbf_t *p) { return p->shift; }
int64_t orc_pbf_n(const orc_pbf_t *p) { return p->n; }

static int is_subset(const orc_pbf_t *p) { return p->n_sub > 0 && p->n_sub < p->m; }

/* pbwt.c:313-337 */
const uint8_t **orc_pbf_read(orc_pbf_t *p)
{
	uint8_t t = 0;
	int g;
	if (rd(p, &t, 1) < 0) return 0;
	if (t == 'S') { /* pbwt.c:319-323: snapshots reload the FULL codec state only */
		for (g = 0; g < p->g; ++g) rd(p, p->pc[g]->cur, (size_t)p->m * 4);
		if (rd(p, &t, 1) < 0) return 0;
	}
	if (t != 'B') return 0;
	for (g = 0; g < p->g; ++g) {
		int32_t l;
		rd(p, &l, 4);
		if (l < 0 || l > p->m) return 0;
		rd(p, p->tmp, l);
		p->tmp[l] = 0;
		if (is_subset(p)) orc_subset_dec(p->m, p->n_sub, p->sub[g], p->tmp, p->pc[g]->u);
		else orc_codec_dec(p->pc[g], p->tmp);
	}
	++p->k;
	return p->ret;
}

/* pbwt.c:374-388 */
int orc_pbf_subset(orc_pbf_t *p, int n_sub, const int *cols)
{
	int g, i;
	if (n_sub <= 0 || n_sub >= p->m || cols == 0) n_sub = 0;
	p->n_sub = n_sub;
	if (n_sub == 0) return 0;
	p->cols = (int*)realloc(p->cols, (size_t)n_sub * sizeof(int));
	memcpy(p->cols, cols, (size_t)n_sub * sizeof(int));
	for (g = 0; g < p->g; ++g) {
		p->sub[g] = (orc_sub_t*)realloc(p->sub[g], (size_t)n_sub * sizeof(orc_sub_t));
		for (i = 0; i < n_sub; ++i) p->sub[g][i].i = i;
		orc_subset_fill(p->m, p->pc[g]->cur, n_sub, p->sub[g], p->cols);
	}
	return 0;
}

/* pbwt.c:349-372 */
int orc_pbf_seek(orc_pbf_t *p, int64_t k)
{
	int g;
	int64_t i, x;
	uint8_t t;
	if (k == p->k) return 0;
	if (k > p->k && k - p->k <= (1LL << p->shift)) {
		while (p->k < k) if (!orc_pbf_read(p)) return -1;
		return 0;
	}
	if (p->idx == 0 || k >= p->n || k < 0) return -1;
	p->pos = p->idx[k >> p->shift];
	rd(p, &t, 1);
	if (t != 'S') return -2; /* the reference asserts (pbwt.c:362) */
	for (g = 0; g < p->g; ++g) {
		rd(p, p->pc[g]->cur, (size_t)p->m * 4);
		if (is_subset(p)) orc_subset_fill(p->m, p->pc[g]->cur, p->n_sub, p->sub[g], p->cols);
	}
	p->k = k >> p->shift << p->shift;
	x = k & ((1LL << p->shift) - 1);
	for (i = 0; i < x; ++i) if (!orc_pbf_read(p)) return -1;
	return 0;
}

/* Algorithmic input bytes of rows [row_beg,row_end): the 'B' records, plus the 'S' records of the
 * checkpoints inside the range when with_snapshots (SURVEY 8d / BASELINE.md "algorithmic bytes"). */
int64_t orc_pbf_row_bytes(orc_pbf_t *p, int64_t row_beg, int64_t row_end, int with_snapshots)
{
	int64_t bytes = 0, k;
	size_t pos;
	if (!p->idx || row_beg < 0 || row_end > p->n || row_beg >= row_end) return 0;
	k = row_beg >> p->shift << p->shift;
	pos = p->idx[k >> p->shift];
	for (; k < row_end; ++k) {
		int g;
		size_t start = pos;
		if (p->buf[pos] == 'S') {
			size_t ssz = 1 + (size_t)p->g * p->m * 4;
			if (with_snapshots && k >= row_beg) bytes += ssz;
			pos += ssz; start = pos;
		}
		++pos; /* 'B' */
		for (g = 0; g < p->g; ++g) { int32_t l; memcpy(&l, p->buf + pos, 4); pos += 4 + l; }
		if (k >= row_beg) bytes += pos - start;
	}
	return bytes;
}

/* ------------------------------------------------------------------ PBF writer */

struct orc_pbfw_s {
	int32_t m, g, shift;
	int64_t n;
	orc_codec_t **pc;
	uint8_t *buf; size_t len, cap;
	uint64_t *idx; int32_t n_idx, m_idx;
};

static void wr(orc_pbfw_t *w, const void *src, size_t n)
{
	if (w->len + n > w->cap) {
		while (w->len + n > w->cap) w->cap = w->cap ? w->cap * 2 : 1 << 16;
		w->buf = (uint8_t*)realloc(w->buf, w->cap);
	}
	memcpy(w->buf + w->len, src, n); w->len += n;
}

/* pbwt.c:199-219 */
orc_pbfw_t *orc_pbfw_new(int m, int g, int shift)
{
	orc_pbfw_t *w = (orc_pbfw_t*)calloc(1, sizeof(*w));
	int32_t v[3];
	int i;
	w->m = m; w->g = g; w->shift = shift;
	w->pc = (orc_codec_t**)calloc(g, sizeof(void*));
	for (i = 0; i < g; ++i) w->pc[i] = orc_codec_new(m);
	v[0] = m; v[1] = g; v[2] = shift;
	wr(w, "PBF\1", 4); wr(w, v, 12);
	return w;
}

/* pbwt.c:292-301: before rows 0, 2^shift, ... dump the running permutation of every plane */
static void wr_checkpoint(orc_pbfw_t *w)
{
	int g;
	if (w->n & ((1LL << w->shift) - 1)) return;
	if (w->n_idx == w->m_idx) {
		w->m_idx = w->m_idx ? w->m_idx * 2 : 8;
		w->idx = (uint64_t*)realloc(w->idx, (size_t)w->m_idx * 8);
	}
	w->idx[w->n_idx++] = w->len;
	wr(w, "S", 1);
	for (g = 0; g < w->g; ++g) wr(w, w->pc[g]->cur, (size_t)w->m * 4);
}

/* pbwt.c:288-311 */
int orc_pbfw_write(orc_pbfw_t *w, const uint8_t *const *a)
{
	int g;
	wr_checkpoint(w);
	wr(w, "B", 1);
	for (g = 0; g < w->g; ++g) {
		orc_codec_enc(w->pc[g], a[g]);
		wr(w, &w->pc[g]->l, 4);
		wr(w, w->pc[g]->u, w->pc[g]->l);
	}
	++w->n;
	return 0;
}

int orc_pbfw_write_rle(orc_pbfw_t *w, const uint8_t *const *rle, const int32_t *l)
{
	int g;
	uint8_t *z = (uint8_t*)malloc((size_t)w->m + 2);
	wr_checkpoint(w);
	wr(w, "B", 1);
	for (g = 0; g < w->g; ++g) {
		wr(w, &l[g], 4);
		wr(w, rle[g], l[g]);
		memcpy(z, rle[g], l[g]); z[l[g]] = 0;
		orc_codec_dec(w->pc[g], z); /* keep the running permutation truthful */
	}
	free(z);
	++w->n;
	return 0;
}

/* pbwt.c:268-276 */
size_t orc_pbfw_finish(orc_pbfw_t *w, uint8_t **buf)
{
	uint64_t off = w->len;
	wr(w, "I", 1);
	wr(w, &w->n, 8);
	wr(w, &w->n_idx, 4);
	wr(w, w->idx, (size_t)w->n_idx * 8);
	wr(w, &off, 8);
	*buf = w->buf;
	return w->len;
}

void orc_pbfw_free(orc_pbfw_t *w)
{
	int i;
	if (!w) return;
	for (i = 0; i < w->g; ++i) orc_codec_free(w->pc[i]);
	free(w->pc); free(w->buf); free(w->idx); free(w);
}
