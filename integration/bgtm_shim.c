/*
 * bgtm_shim.c -- seam B: bgtm_read (bgt.c:880-888) served from device-computed results.
 *
 * This file is what a maintainer of the reference adds to libbgt.  It is compiled against
 * the reference's public headers; the reference's bgt.c is compiled unchanged except that three entry points
 * are renamed on the command line (-Dbgtm_read=ref_bgtm_read -Dbgtm_set_flt_site=ref_bgtm_set_flt_site
 * -Dbgtm_reader_destroy=ref_bgtm_reader_destroy) so that the definitions below take their place.
 *
 * Per record the reference does (bgt.c:797-878): read the site (bgt_read_core), pbf_seek+pbf_read (CPU PBWT
 * decode), copy the two decoded planes, bgtm_cal_info (the 65-91 % hotspot), bgtm_fill_info, the -f verdict.
 * Here sites are pulled ahead in batches, ONE b200_scan per batch produces AC/AN, verdicts and (if genotypes are
 * printed) the decoded planes for all of them, and records are then assembled with the reference's own
 * bcfcpy_min / bgtm_fill_info / bgt_gen_gt so that the output bytes are identical.
 *
 * Queries outside the accelerated path (several BGT files, -a/-S/-H allele queries) are passed to the
 * reference's own bgtm_read, whose row decode then goes through seam A (pbwt_shim.c), i.e. still the GPU.
 *
 * Library code: nothing here terminates the process.  A device failure makes bgtm_read return -2 (the reference's
 * only negative value is -1 = end of data, bgt.c:880-888) with the message on stderr and in pbf_b200_strerror().
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bgt.h"
#include "../include/bgt_b200.h"

int ref_bgtm_read(bgtm_t *bm, bcf1_t *b);
int ref_bgtm_set_flt_site(bgtm_t *bm, const char *expr);
void ref_bgtm_reader_destroy(bgtm_t *bm);
int bgt_read_core(bgt_t *bgt);                                   /* bgt.c:315-331 */
void bgtm_fill_info(const bcf_hdr_t *h, const bgt_info_t *ss, bcf1_t *b); /* bgt.c:721-733 */
int bgtm_gen_tbl_line(bgtm_t *bm, const bgt_info_t *ss, const bcf1_t *b);  /* bgt.c:775-795 */
int bgtm_pass_site_flt(const bgt_info_t *ss, kexpr_t *flt);      /* bgt.c:712-719 */
void bgt_gen_gt(const bcf_hdr_t *h, bcf1_t *b, int m, const uint8_t **a, int32_t *mgs); /* bgt.c:290-313 */
b200_ctx_t *pbf_b200_ctx(void);                                  /* pbwt_shim.c */
const uint8_t *pbf_b200_image(const pbf_t *pb, size_t *len);
void pbf_b200_route_add(int slot, int64_t n);

typedef struct accel_s {
	struct accel_s *next;
	bgtm_t *bm;
	char *flt;
	int decided, eligible, host_flt, failed;
	b200_pbf_t *win; int64_t win_beg, win_end;
	b200_query_t *q;
	int n_rec, cur, cap, eof, has_pending;
	bcf1_t **rec; int64_t *rows;
	int64_t *ridx;                    /* per record: index of its row in the batch's outputs (regions back to back) */
	int64_t *rg_beg, *rg_cnt; int rg_cap;   /* the batch's regions: runs of records whose rows lie close together */
	size_t out_cap_rows;
	int32_t *counts; uint8_t *pass, *hap[2];
	int stride, n_track;
} accel_t;

static accel_t *g_accel;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
/* bgtm_t is ABI (the Go server allocates it itself, bgt-server.go:22-30), so the per-reader state lives in a side table.
 * The table is consulted once per reader and thread, not once per record: every thread remembers the reader it served
 * last (a server thread works on one bgtm_t at a time); g_gen invalidates the memo when any reader is destroyed. */
static volatile unsigned g_gen = 1;
static __thread bgtm_t *t_bm;
static __thread accel_t *t_accel;
static __thread unsigned t_gen;

static accel_t *accel_get(bgtm_t *bm, int create)
{
	accel_t *a;
	if (t_bm == bm && t_gen == g_gen && t_accel) return t_accel;
	pthread_mutex_lock(&g_lock);
	for (a = g_accel; a; a = a->next) if (a->bm == bm) break;
	if (a == 0 && create) {
		a = (accel_t*)calloc(1, sizeof(accel_t));
		a->bm = bm; a->next = g_accel; g_accel = a;
	}
	t_bm = bm; t_accel = a; t_gen = g_gen;
	pthread_mutex_unlock(&g_lock);
	return a;
}

static void accel_drop(bgtm_t *bm)
{
	accel_t **pp, *a = 0;
	int i;
	pthread_mutex_lock(&g_lock);
	for (pp = &g_accel; *pp; pp = &(*pp)->next) if ((*pp)->bm == bm) { a = *pp; *pp = a->next; break; }
	++g_gen;
	pthread_mutex_unlock(&g_lock);
	if (a == 0) return;
	if (a->q) b200_query_destroy(a->q);
	if (a->win) b200_pbf_close(a->win);
	for (i = 0; i < a->cap; ++i) if (a->rec[i]) bcf_destroy1(a->rec[i]);
	free(a->rec); free(a->rows); free(a->ridx); free(a->rg_beg); free(a->rg_cnt); free(a->flt);
	b200_host_free(a->counts); b200_host_free(a->pass); b200_host_free(a->hap[0]); b200_host_free(a->hap[1]);
	free(a);
}

int bgtm_set_flt_site(bgtm_t *bm, const char *expr)              /* bgt.c:444-455 */
{
	accel_t *a = accel_get(bm, 1);
	free(a->flt);
	a->flt = expr ? strdup(expr) : 0;
	/* a filter set or changed after the first read: the device query (compiled filter) and the routing decision are stale */
	if (a->q) { b200_query_destroy(a->q); a->q = 0; }
	a->decided = 0; a->host_flt = 0;
	return ref_bgtm_set_flt_site(bm, expr);
}

void bgtm_reader_destroy(bgtm_t *bm)                             /* bgt.c:376-403 */
{
	accel_drop(bm);
	ref_bgtm_reader_destroy(bm);
}

/* a device failure: reported, remembered (the reader stays failed), never fatal to the host process and never papered over
 * by the CPU path */
static int fail(accel_t *a, const char *what)
{
	fprintf(stderr, "[E::bgt_b200] %s: %s\n", what, b200_strerror());
	a->failed = 1; a->n_rec = a->cur = 0;
	return -1;
}

static int need_ac(const bgtm_t *bm) /* bgt.c:850 */
{
	return (bm->flag & BGT_F_SET_AC) || bm->site_flt || bm->n_fields > 0 || bm->n_groups > 1;
}

static void decide(accel_t *a)
{
	bgtm_t *bm = a->bm;
	const char *off = getenv("BGT_B200_DISABLE");
	a->decided = 1;
	a->eligible = bm->n_bgt == 1 && bm->h_al == 0 && !(bm->flag & (BGT_F_CNT_AL | BGT_F_CNT_HAP)) &&
	              bm->bgt[0]->n_out > 0 && !(off && *off == '1');
}

/* pull the next batch of site records and run the GPU scan over their row range; 0 or -1 (device failure) */
static int fill_batch(accel_t *a)
{
	bgtm_t *bm = a->bm;
	bgt_t *bgt = bm->bgt[0];
	b200_ctx_t *ctx = pbf_b200_ctx();
	const int want_gt = !(bm->flag & BGT_F_NO_GT);
	const int n_track = bgt->n_out << 1;
	size_t map_len;
	const uint8_t *map = pbf_b200_image(bgt->pb, &map_len);
	int64_t rows_cap, covered = 0, n_rows;
	b200_scan_out_t so;
	unsigned flags = 0;
	int n_rg = 0, i;
	const int64_t gap_max = 64;       /* rows between two records that are still scanned through rather than split into two regions */
	if (ctx == 0) { a->failed = 1; return -1; }
	if (map == 0) { fprintf(stderr, "[E::bgt_b200] the PBF handle was not opened by the B200 seam\n"); a->failed = 1; return -1; }
	/* batch geometry: bounded by the decoded-plane bytes when genotypes are printed */
	rows_cap = want_gt ? (64LL << 20) / (n_track > 0 ? n_track : 1) : 65536;
	if (rows_cap < 1) rows_cap = 1;
	if (rows_cap > 65536) rows_cap = 65536;
	if (a->cap < rows_cap + 1) {
		int old = a->cap;
		a->cap = (int)rows_cap + 1;
		a->rec = (bcf1_t**)realloc(a->rec, a->cap * sizeof(void*));
		a->rows = (int64_t*)realloc(a->rows, a->cap * sizeof(int64_t));
		a->ridx = (int64_t*)realloc(a->ridx, a->cap * sizeof(int64_t));
		for (i = old; i < a->cap; ++i) a->rec[i] = bcf_init1();
	}
	/* the record that closed the previous batch opens this one */
	if (a->has_pending) {
		bcf1_t *t = a->rec[0]; a->rec[0] = a->rec[a->n_rec]; a->rec[a->n_rec] = t;
		a->rows[0] = a->rows[a->n_rec];
		a->n_rec = 1; a->has_pending = 0;
	} else a->n_rec = 0;
	a->cur = 0;
	while (!a->eof && a->n_rec < rows_cap) {
		int row = bgt_read_core(bgt);                            /* region / BED handling stays the reference's */
		if (row < 0) { a->eof = 1; break; }
		bcfcpy(a->rec[a->n_rec], bgt->b0);
		a->rows[a->n_rec] = row;
		if (a->n_rec > 0) {
			/* rows the batch covers: records that lie close together are scanned through, a far jump (`-B` regions, sparse
			 * site lists) starts a new region of the batch instead of dragging every row in between along */
			const int64_t first = a->rows[0], prev = a->rows[a->n_rec - 1];
			const int64_t wend = a->win && first >= a->win_beg && first < a->win_end ? a->win_end : -1;
			const int64_t step = row - prev > gap_max ? 1 : row - prev;
			const int64_t far = row - first >= (1LL << 30) / (8LL * pbf_get_m(bgt->pb) + 1) * (1LL << pbf_get_shift(bgt->pb));   /* beyond one window */
			if (covered + step >= rows_cap || (wend >= 0 && row >= wend) || (wend < 0 && far) || row < prev) { a->has_pending = 1; break; }
			covered += step;
		} else covered = 1;
		++a->n_rec;
	}
	if (a->n_rec == 0) return 0;
	/* residency: the checkpoint blocks around the batch */
	if (a->win == 0 || a->rows[0] < a->win_beg || a->rows[a->n_rec - 1] >= a->win_end) {
		const int shift = pbf_get_shift(bgt->pb);
		const int64_t BS = 1LL << shift, m = pbf_get_m(bgt->pb);
		int64_t wblocks = (1LL << 30) / (8 * m + 256 * BS), beg, end;
		if (wblocks < 1) wblocks = 1;
		beg = a->rows[0] / BS * BS;
		end = beg + wblocks * BS;
		if (end <= a->rows[a->n_rec - 1]) end = (a->rows[a->n_rec - 1] / BS + 1) * BS;
		if (a->q) { b200_query_destroy(a->q); a->q = 0; }
		if (a->win) b200_pbf_close(a->win);
		a->win = b200_pbf_load(ctx, map, map_len, beg, end);
		if (a->win == 0) return fail(a, "loading the PBF window");
		a->win_beg = b200_pbf_row_beg(a->win); a->win_end = b200_pbf_row_end(a->win);
	}
	if (a->q == 0) {
		int err = 0;
		a->host_flt = 0;
		a->q = b200_query_create(ctx, a->win, bgt->n_out, bgt->out, bm->group, bm->n_groups, a->flt, &err);
		if (a->q == 0 && err && b200_errcode() == B200_E_FILTER_SYNTAX) { /* kexpr accepted it but the device compiler did not: verdict on the host from device counts */
			a->host_flt = 1;
			a->q = b200_query_create(ctx, a->win, bgt->n_out, bgt->out, bm->group, bm->n_groups, 0, &err);
		}
		if (a->q == 0) return fail(a, "preparing the query");
		a->stride = b200_query_counts_stride(a->q);
		a->n_track = b200_query_n_track(a->q);
	}
	/* the batch's regions and every record's place in the outputs */
	for (i = 0, n_rows = 0; i < a->n_rec; ++i) {
		if (i == 0 || a->rows[i] - a->rows[i - 1] > gap_max) {
			if (n_rg == a->rg_cap) {
				a->rg_cap = a->rg_cap ? a->rg_cap << 1 : 64;
				a->rg_beg = (int64_t*)realloc(a->rg_beg, a->rg_cap * sizeof(int64_t));
				a->rg_cnt = (int64_t*)realloc(a->rg_cnt, a->rg_cap * sizeof(int64_t));
			}
			if (n_rg) n_rows += a->rg_cnt[n_rg - 1];
			a->rg_beg[n_rg] = a->rows[i]; a->rg_cnt[n_rg] = 0; ++n_rg;
		}
		a->rg_cnt[n_rg - 1] = a->rows[i] - a->rg_beg[n_rg - 1] + 1;
		a->ridx[i] = n_rows + (a->rows[i] - a->rg_beg[n_rg - 1]);
	}
	n_rows += a->rg_cnt[n_rg - 1];
	if ((size_t)n_rows > a->out_cap_rows) {
		b200_host_free(a->counts); b200_host_free(a->pass); b200_host_free(a->hap[0]); b200_host_free(a->hap[1]);
		a->out_cap_rows = (size_t)n_rows;
		a->counts = (int32_t*)b200_host_alloc(a->out_cap_rows * a->stride * sizeof(int32_t));
		a->pass = (uint8_t*)b200_host_alloc(a->out_cap_rows);
		a->hap[0] = a->hap[1] = 0;
		if (want_gt) {
			a->hap[0] = (uint8_t*)b200_host_alloc(a->out_cap_rows * a->n_track);
			a->hap[1] = (uint8_t*)b200_host_alloc(a->out_cap_rows * a->n_track);
		}
	} else if (want_gt && a->hap[0] == 0) {
		a->hap[0] = (uint8_t*)b200_host_alloc(a->out_cap_rows * a->n_track);
		a->hap[1] = (uint8_t*)b200_host_alloc(a->out_cap_rows * a->n_track);
	}
	memset(&so, 0, sizeof(so));
	so.counts = a->counts; so.pass = a->pass; flags |= B200_SCAN_COUNTS;
	if (want_gt) { so.hap_bytes[0] = a->hap[0]; so.hap_bytes[1] = a->hap[1]; flags |= B200_SCAN_HAP_BYTES; }
	if (b200_scan_regions(ctx, a->win, a->q, n_rg, a->rg_beg, a->rg_cnt, flags, &so) != n_rows) return fail(a, "scan");
	pbf_b200_route_add(2, 1);
	if (n_rg > 1) pbf_b200_route_add(7, 1);
	return 0;
}

int bgtm_read(bgtm_t *bm, bcf1_t *b)                             /* bgt.c:880-888 */
{
	accel_t *a;
	bgt_t *bgt;
	if (bm->h_out == 0) bgtm_prepare(bm);
	a = accel_get(bm, 1);
	if (!a->decided) decide(a);
	if (!a->eligible) { pbf_b200_route_add(3, 1); return ref_bgtm_read(bm, b); }
	if (a->failed) return -2;
	bgt = bm->bgt[0];
	for (;;) {
		const bcf1_t *b0;
		const int32_t *c;
		int64_t rr;
		int l_ref;
		if (a->cur >= a->n_rec) {
			if (a->eof && !a->has_pending) return -1;
			if (fill_batch(a) != 0) return -2;
			if (a->n_rec == 0) return -1;
		}
		b0 = a->rec[a->cur];
		rr = a->ridx[a->cur];
		++a->cur;
		bm->n_gt_read += bgt->n_out;                             /* bgt.c:807 */
		l_ref = bcfcpy_min(b, b0, b0->n_allele > 2 ? "<M>" : 0);  /* bgt.c:823 */
		if (l_ref != b->rlen) {                                  /* bgt.c:824-827 */
			int32_t val = b->pos + b->rlen;
			bcf_append_info_ints(bm->h_out, b, "END", 1, &val);
		}
		if (!(bm->flag & BGT_F_NO_GT)) {                         /* bgt.c:835-836, from the device-decoded planes */
			memcpy(bm->a[0], a->hap[0] + (size_t)rr * a->n_track, a->n_track);
			memcpy(bm->a[1], a->hap[1] + (size_t)rr * a->n_track, a->n_track);
		}
		if (need_ac(bm)) {                                       /* bgt.c:850-857 with device-computed counts */
			bgt_info_t ss;
			int i, pass;
			c = a->counts + (size_t)rr * a->stride;
			memset(&ss, 0, sizeof(ss));
			ss.an = c[0]; ss.ac[0] = c[1]; ss.ac[1] = c[2]; ss.n_groups = bm->n_groups;
			for (i = 0; i < bm->n_groups; ++i)
				ss.gan[i] = c[3 + 3 * i], ss.gac[i][0] = c[4 + 3 * i], ss.gac[i][1] = c[5 + 3 * i];
			bgtm_fill_info(bm->h_out, &ss, b);
			if (bm->n_fields > 0) bgtm_gen_tbl_line(bm, &ss, b);
			pass = a->host_flt ? bgtm_pass_site_flt(&ss, bm->site_flt) : a->pass[rr];
			if (bm->site_flt && !pass) continue;
		}
		if (!(bm->flag & BGT_F_NO_GT))
			bgt_gen_gt(bm->h_out, b, bm->n_out, (const uint8_t**)bm->a, bm->mgs); /* bgt.c:885-886 */
		return 0;
	}
}
