/*
 * bgtm_shim.c -- seam B: bgtm_read (bgt.c:880-888) served from device-computed results.
 *
 * This file is what a maintainer of the reference adds to libbgt.  It is compiled against
 * the reference's public headers; the reference's bgt.c is compiled unchanged except that three entry points
 * are renamed on the command line (-Dbgtm_read=ref_bgtm_read -Dbgtm_set_flt_site=ref_bgtm_set_flt_site
 * -Dbgtm_reader_destroy=ref_bgtm_reader_destroy) so that the definitions below take their place.
 *
 * Per record the reference does (bgt.c:797-878): read the next site of every file (bgt_read_rec: bgt_read_core +
 * pbf_seek + pbf_read = CPU PBWT decode), pick the smallest site (bcfcmp), copy the decoded planes of the files that
 * hold it -- missing calls for the files that do not --, bgtm_cal_info (the 65-91 % hotspot), bgtm_fill_info, the -f
 * verdict.  Here every file has a FEEDER: its sites are pulled ahead in batches through the reference's own
 * bgt_read_core (region / BED / -i logic stays the reference's), ONE b200_scan_regions per batch produces AC/AN per
 * group, verdicts and (if genotypes are printed) the decoded planes for all of them.  Records are then merged and
 * assembled with the reference's own bcfcmp / bcfcpy_min / bgtm_fill_info / bgt_gen_gt, so that the output bytes are
 * identical.  With several files the per-group counts of the files that hold the site are added up (a file without
 * the site contributes missing calls, i.e. nothing to AN/AC -- bgt.c:837-840, 746-756) and the -f verdict is taken on
 * the sums with the reference's own evaluator, bgtm_pass_site_flt.
 *
 * Allele queries (-a, with -S / -H: bgt.c:844-848, 859-876): the reference's bgt_read_core already drops the sites that carry none
 * of the listed alleles, so a feeder's batch is a handful of scattered rows (one region each); the per-sample bookkeeping of -S/-H
 * runs on the device-decoded planes exactly where the reference runs it on pbf_read's.
 *
 * Library code: nothing here terminates the process.  A device failure makes bgtm_read return -2 (the reference's
 * only negative value is -1 = end of data, bgt.c:880-888) with the message on stderr and in pbf_b200_strerror().
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bgt.h"
#include "khash.h"
#include "../include/bgt_b200.h"

KHASH_SET_INIT_STR(str)                                          /* the same instantiation as bgt.c:12: bm->h_al points to one of these */

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

/* one BGT file of the reader: its resident window, its query, the batch of sites pulled ahead and their device results */
typedef struct {
	bgt_t *bgt;
	int off;                          /* first selected sample of this file among the reader's (order of bgtm_prepare, bgt.c:613-620) */
	b200_pbf_t *win; int64_t win_beg, win_end;
	b200_query_t *q;
	int n_rec, cur, cap, eof, has_pending;
	bcf1_t **rec; int64_t *rows;
	int64_t *ridx;                    /* per record: index of its row in the batch's outputs (regions back to back) */
	int64_t *rg_beg, *rg_cnt; int rg_cap;   /* the batch's regions: runs of records whose rows lie close together */
	size_t out_cap_rows;
	int32_t *counts; uint8_t *pass, *hap[2];
	int stride, n_track;
} feeder_t;

typedef struct accel_s {
	struct accel_s *next;
	bgtm_t *bm;
	char *flt;
	int decided, eligible, host_flt, failed;
	int n_f;
	feeder_t *f;
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

static void feeder_release(feeder_t *f)
{
	int i;
	if (f->q) b200_query_destroy(f->q);
	if (f->win) b200_pbf_close(f->win);
	for (i = 0; i < f->cap; ++i) if (f->rec[i]) bcf_destroy1(f->rec[i]);
	free(f->rec); free(f->rows); free(f->ridx); free(f->rg_beg); free(f->rg_cnt);
	b200_host_free(f->counts); b200_host_free(f->pass); b200_host_free(f->hap[0]); b200_host_free(f->hap[1]);
	memset(f, 0, sizeof(*f));
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
	for (i = 0; i < a->n_f; ++i) feeder_release(&a->f[i]);
	free(a->f); free(a->flt);
	free(a);
}

int bgtm_set_flt_site(bgtm_t *bm, const char *expr)              /* bgt.c:444-455 */
{
	accel_t *a = accel_get(bm, 1);
	int i;
	free(a->flt);
	a->flt = expr ? strdup(expr) : 0;
	/* a filter set or changed after the first read: the device queries (compiled filter) and the routing decision are stale */
	for (i = 0; i < a->n_f; ++i) if (a->f[i].q) { b200_query_destroy(a->f[i].q); a->f[i].q = 0; }
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
	a->failed = 1;
	return -1;
}

static int need_ac(const bgtm_t *bm) /* bgt.c:850 */
{
	return (bm->flag & BGT_F_SET_AC) || bm->site_flt || bm->n_fields > 0 || bm->n_groups > 1;
}

/* which of the listed alleles a site carries (bgt.c:252-270): 1 its ALT, 2 its REF, 0 neither */
static int allele_listed(void *h_al, const bcf_hdr_t *hdr, const bcf1_t *b)
{
	khash_t(str) *h = (khash_t(str)*)h_al;
	bgt_allele_t alt, ref;
	kstring_t key = {0,0,0};
	int ret = 0;
	memset(&alt, 0, sizeof(alt)); memset(&ref, 0, sizeof(ref));
	bgt_al_from_bcf(hdr, b, &alt, &ref);
	bgt_al_format(&alt, &key);
	if (kh_get(str, h, key.s) != kh_end(h)) ret = 1;
	else {
		bgt_al_format(&ref, &key);
		if (kh_get(str, h, key.s) != kh_end(h)) ret = 2;
	}
	free(key.s); free(alt.chr.s); free(ref.chr.s);
	return ret;
}

/* the decoded planes are wanted for the genotype columns and for the per-sample bookkeeping of -S / -H */
static int want_planes(const bgtm_t *bm)
{
	return !(bm->flag & BGT_F_NO_GT) || (bm->h_al && (bm->flag & (BGT_F_CNT_AL | BGT_F_CNT_HAP)));
}

static void decide(accel_t *a)
{
	bgtm_t *bm = a->bm;
	const char *off = getenv("BGT_B200_DISABLE");
	int i, o;
	a->decided = 1;
	a->eligible = bm->n_bgt >= 1 && bm->n_out > 0 && !(off && *off == '1');
	if (!a->eligible) return;
	/* several files: the verdict is taken on the counts summed over the files, by the reference's evaluator on the host */
	a->host_flt = bm->n_bgt > 1;
	if (a->n_f == bm->n_bgt) return;
	a->n_f = bm->n_bgt;
	a->f = (feeder_t*)calloc(a->n_f, sizeof(feeder_t));
	for (i = o = 0; i < a->n_f; ++i) {
		a->f[i].bgt = bm->bgt[i]; a->f[i].off = o;
		o += bm->bgt[i]->n_out;
		if (bm->bgt[i]->n_out == 0) a->f[i].eof = 1;             /* bgt_read_rec never reads a file without selected samples (bgt.c:338) */
	}
}

/* pull the next batch of site records of one file and run the GPU scan over their rows; 0 or -1 (device failure) */
static int fill_batch(accel_t *a, feeder_t *f)
{
	bgtm_t *bm = a->bm;
	bgt_t *bgt = f->bgt;
	b200_ctx_t *ctx = pbf_b200_ctx();
	const int want_gt = want_planes(bm);
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
	if (f->cap < rows_cap + 1) {
		int old = f->cap;
		f->cap = (int)rows_cap + 1;
		f->rec = (bcf1_t**)realloc(f->rec, f->cap * sizeof(void*));
		f->rows = (int64_t*)realloc(f->rows, f->cap * sizeof(int64_t));
		f->ridx = (int64_t*)realloc(f->ridx, f->cap * sizeof(int64_t));
		for (i = old; i < f->cap; ++i) f->rec[i] = bcf_init1();
	}
	/* the record that closed the previous batch opens this one */
	if (f->has_pending) {
		bcf1_t *t = f->rec[0]; f->rec[0] = f->rec[f->n_rec]; f->rec[f->n_rec] = t;
		f->rows[0] = f->rows[f->n_rec];
		f->n_rec = 1; f->has_pending = 0;
	} else f->n_rec = 0;
	f->cur = 0;
	while (!f->eof && f->n_rec < rows_cap) {
		int row = bgt_read_core(bgt);                            /* region / BED handling stays the reference's */
		if (row < 0) { f->eof = 1; break; }
		bcfcpy(f->rec[f->n_rec], bgt->b0);
		f->rows[f->n_rec] = row;
		if (f->n_rec > 0) {
			/* rows the batch covers: records that lie close together are scanned through, a far jump (`-B` regions, sparse
			 * site lists) starts a new region of the batch instead of dragging every row in between along */
			const int64_t first = f->rows[0], prev = f->rows[f->n_rec - 1];
			const int64_t wend = f->win && first >= f->win_beg && first < f->win_end ? f->win_end : -1;
			const int64_t step = row - prev > gap_max ? 1 : row - prev;
			const int64_t far = row - first >= (1LL << 30) / (8LL * pbf_get_m(bgt->pb) + 1) * (1LL << pbf_get_shift(bgt->pb));   /* beyond one window */
			if (covered + step >= rows_cap || (wend >= 0 && row >= wend) || (wend < 0 && far) || row < prev) { f->has_pending = 1; break; }
			covered += step;
		} else covered = 1;
		++f->n_rec;
	}
	if (f->n_rec == 0) return 0;
	/* residency: the checkpoint blocks around the batch */
	if (f->win == 0 || f->rows[0] < f->win_beg || f->rows[f->n_rec - 1] >= f->win_end) {
		const int shift = pbf_get_shift(bgt->pb);
		const int64_t BS = 1LL << shift, m = pbf_get_m(bgt->pb);
		int64_t wblocks = (1LL << 30) / (8 * m + 256 * BS), beg, end;
		if (wblocks < 1) wblocks = 1;
		beg = f->rows[0] / BS * BS;
		end = beg + wblocks * BS;
		if (end <= f->rows[f->n_rec - 1]) end = (f->rows[f->n_rec - 1] / BS + 1) * BS;
		if (f->q) { b200_query_destroy(f->q); f->q = 0; }
		if (f->win) b200_pbf_close(f->win);
		f->win = b200_pbf_load(ctx, map, map_len, beg, end);
		if (f->win == 0) return fail(a, "loading the PBF window");
		f->win_beg = b200_pbf_row_beg(f->win); f->win_end = b200_pbf_row_end(f->win);
	}
	if (f->q == 0) {
		int err = 0;
		const uint32_t *grp = bm->group + f->off;               /* this file's slice of the reader's group labels (bgt.c:617) */
		if (!a->host_flt) {
			f->q = b200_query_create(ctx, f->win, bgt->n_out, bgt->out, grp, bm->n_groups, a->flt, &err);
			if (f->q == 0 && err && b200_errcode() == B200_E_FILTER_SYNTAX) a->host_flt = 1;   /* kexpr accepted it but the device compiler did not: verdict on the host from device counts */
		}
		if (f->q == 0 && a->host_flt) f->q = b200_query_create(ctx, f->win, bgt->n_out, bgt->out, grp, bm->n_groups, 0, &err);
		if (f->q == 0) return fail(a, "preparing the query");
		f->stride = b200_query_counts_stride(f->q);
		f->n_track = b200_query_n_track(f->q);
	}
	/* the batch's regions and every record's place in the outputs */
	for (i = 0, n_rows = 0; i < f->n_rec; ++i) {
		if (i == 0 || f->rows[i] - f->rows[i - 1] > gap_max) {
			if (n_rg == f->rg_cap) {
				f->rg_cap = f->rg_cap ? f->rg_cap << 1 : 64;
				f->rg_beg = (int64_t*)realloc(f->rg_beg, f->rg_cap * sizeof(int64_t));
				f->rg_cnt = (int64_t*)realloc(f->rg_cnt, f->rg_cap * sizeof(int64_t));
			}
			if (n_rg) n_rows += f->rg_cnt[n_rg - 1];
			f->rg_beg[n_rg] = f->rows[i]; f->rg_cnt[n_rg] = 0; ++n_rg;
		}
		f->rg_cnt[n_rg - 1] = f->rows[i] - f->rg_beg[n_rg - 1] + 1;
		f->ridx[i] = n_rows + (f->rows[i] - f->rg_beg[n_rg - 1]);
	}
	n_rows += f->rg_cnt[n_rg - 1];
	if ((size_t)n_rows > f->out_cap_rows) {
		b200_host_free(f->counts); b200_host_free(f->pass); b200_host_free(f->hap[0]); b200_host_free(f->hap[1]);
		f->out_cap_rows = (size_t)n_rows;
		f->counts = (int32_t*)b200_host_alloc(f->out_cap_rows * f->stride * sizeof(int32_t));
		f->pass = (uint8_t*)b200_host_alloc(f->out_cap_rows);
		f->hap[0] = f->hap[1] = 0;
		if (want_gt) {
			f->hap[0] = (uint8_t*)b200_host_alloc(f->out_cap_rows * f->n_track);
			f->hap[1] = (uint8_t*)b200_host_alloc(f->out_cap_rows * f->n_track);
		}
	} else if (want_gt && f->hap[0] == 0) {
		f->hap[0] = (uint8_t*)b200_host_alloc(f->out_cap_rows * f->n_track);
		f->hap[1] = (uint8_t*)b200_host_alloc(f->out_cap_rows * f->n_track);
	}
	memset(&so, 0, sizeof(so));
	so.counts = f->counts; so.pass = f->pass; flags |= B200_SCAN_COUNTS;
	if (want_gt) { so.hap_bytes[0] = f->hap[0]; so.hap_bytes[1] = f->hap[1]; flags |= B200_SCAN_HAP_BYTES; }
	if (b200_scan_regions(ctx, f->win, f->q, n_rg, f->rg_beg, f->rg_cnt, flags, &so) != n_rows) return fail(a, "scan");
	pbf_b200_route_add(2, 1);
	if (n_rg > 1) pbf_b200_route_add(7, 1);
	return 0;
}

int bgtm_read(bgtm_t *bm, bcf1_t *b)                             /* bgt.c:880-888 and bgtm_read_core, bgt.c:797-878 */
{
	accel_t *a;
	int want_gt;
	if (bm->h_out == 0) bgtm_prepare(bm);
	want_gt = want_planes(bm);
	a = accel_get(bm, 1);
	if (!a->decided) decide(a);
	if (!a->eligible) { pbf_b200_route_add(3, 1); return ref_bgtm_read(bm, b); }
	if (a->failed) return -2;
	for (;;) {
		const bcf1_t *b0 = 0;
		int i, g, max_allele = 0, n_rest = 0, l_ref, pass_dev = 1, al_ret = 0;
		bgt_info_t ss;
		/* every file's next site (bgt.c:803-808); a feeder refills its batch when it has run dry */
		for (i = 0; i < a->n_f; ++i) {
			feeder_t *f = &a->f[i];
			if (f->cur >= f->n_rec && !(f->eof && !f->has_pending))
				if (fill_batch(a, f) != 0) return -2;
			if (f->cur < f->n_rec) { ++n_rest; bm->n_gt_read += f->bgt->n_out; }   /* bgt.c:807 */
		}
		if (n_rest == 0) return -1;
		/* the smallest site and the largest allele count among its copies (bgt.c:811-820) */
		for (i = 0; i < a->n_f; ++i) {
			feeder_t *f = &a->f[i];
			const bcf1_t *r;
			if (f->cur >= f->n_rec) continue;
			r = f->rec[f->cur];
			if (b0) {
				const int j = bcfcmp(b0, r);
				if (j > 0) b0 = r, max_allele = r->n_allele;
				else if (j == 0 && r->n_allele > max_allele) max_allele = r->n_allele;
			} else b0 = r, max_allele = r->n_allele;
		}
		l_ref = bcfcpy_min(b, b0, max_allele > 2 ? "<M>" : 0);    /* bgt.c:823 */
		if (l_ref != b->rlen) {                                  /* bgt.c:824-827 */
			int32_t val = b->pos + b->rlen;
			bcf_append_info_ints(bm->h_out, b, "END", 1, &val);
		}
		/* the files that hold this site hand over their device results; the others contribute missing calls (bgt.c:829-842) */
		memset(&ss, 0, sizeof(ss));
		ss.n_groups = bm->n_groups;
		for (i = 0; i < a->n_f; ++i) {
			feeder_t *f = &a->f[i];
			const int n2 = f->bgt->n_out << 1;
			if (n2 == 0) continue;
			if (f->cur < f->n_rec && bcfcmp(b, f->rec[f->cur]) == 0) {
				const int64_t rr = f->ridx[f->cur];
				const int32_t *c = f->counts + (size_t)rr * f->stride;
				++f->cur;
				ss.an += c[0]; ss.ac[0] += c[1]; ss.ac[1] += c[2];
				for (g = 0; g < bm->n_groups; ++g)
					ss.gan[g] += c[3 + 3 * g], ss.gac[g][0] += c[4 + 3 * g], ss.gac[g][1] += c[5 + 3 * g];
				pass_dev = f->pass[rr];
				if (want_gt) {                                       /* bgt.c:835-836, from the device-decoded planes */
					memcpy(bm->a[0] + 2 * f->off, f->hap[0] + (size_t)rr * f->n_track, n2);
					memcpy(bm->a[1] + 2 * f->off, f->hap[1] + (size_t)rr * f->n_track, n2);
				}
			} else if (want_gt) {                                    /* bgt.c:838-839 */
				memset(bm->a[0] + 2 * f->off, 0, n2);
				memset(bm->a[1] + 2 * f->off, 1, n2);
			}
		}
		if (bm->h_al && (al_ret = allele_listed(bm->h_al, bm->h_out, b)) == 0) continue;   /* bgt.c:844-848 */
		if (need_ac(bm)) {                                       /* bgt.c:850-857 with device-computed counts */
			bgtm_fill_info(bm->h_out, &ss, b);
			if (bm->n_fields > 0) bgtm_gen_tbl_line(bm, &ss, b);
			if (bm->site_flt && !(a->host_flt ? bgtm_pass_site_flt(&ss, bm->site_flt) : pass_dev)) continue;
		}
		if (bm->h_al) {                                          /* bgt.c:859-876 on the device-decoded planes */
			const uint8_t *a0 = bm->a[0], *a1 = bm->a[1];
			if ((bm->flag & BGT_F_CNT_AL) && bm->alcnt) {            /* -S: one more listed allele seen in these samples */
				const int code = al_ret == 2 ? 0 : 1;                /* the listed allele is the site's REF (code 0) or its ALT (code 1) */
				for (i = 0; i < bm->n_out; ++i)
					bm->alcnt[i] += (a0[2 * i] | a1[2 * i] << 1) == code || (a0[2 * i + 1] | a1[2 * i + 1] << 1) == code;
			}
			if ((bm->flag & BGT_F_CNT_HAP) && bm->hap)                /* -H: bit n_aal of every haplotype that carries the ALT */
				for (i = 0; i < bm->n_out << 1; ++i)
					if (a0[i] == 1 && a1[i] == 0) bm->hap[i] |= 1ULL << bm->n_aal;
			bgt_al_from_bcf(bm->h_out, b, &bm->aal[bm->n_aal++], 0);
		}
		if (!(bm->flag & BGT_F_NO_GT)) bgt_gen_gt(bm->h_out, b, bm->n_out, (const uint8_t**)bm->a, bm->mgs); /* bgt.c:885-886 */
		return 0;
	}
}
