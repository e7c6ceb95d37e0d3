/*
 * view_fast.c -- `bgt view` with the whole per-site pipeline on the device, on one or several GPUs (8e).
 *
 * The reference's view.c is compiled unchanged with -Dmain_view=ref_main_view; this main_view looks at the options
 * first.  For the VCF scan of one BGT -- `view [-G] [-C] [-f EXPR] [-s EXPR ...] prefix`, BASELINE configs 2-5 -- set-up
 * and the header are the reference's own calls (bgt_open, bgtm_reader_init, bgtm_set_flag,
 * bgtm_set_flt_site, bgtm_add_group, bgtm_prepare, vcf_hdr_write: view.c:99-147), and the record loop of
 * view.c:150-155 (bgtm_read + vcf_write1 per site on the host thread) is replaced by the device pipeline: the .bcf/.csi are
 * inflated, indexed and parsed on the GPU (b200_sites_load), the .pbf is scanned (b200_view_text_ex -> b200_scan) and the
 * VCF lines of the passing sites come back as text, window by window when genotype columns are printed.
 *
 * Several GPUs (BGT_B200_DEVICES=0,1,... or "all"): the file is cut into region shards of whole checkpoint blocks
 * (pbwt.c:292-301: every block starts with a snapshot and decodes on its own), one host thread + context per GPU loads
 * and scans its shard, the shard texts are written in shard order behind the single header, and the whole-cohort totals
 * (sum AN, sum AC, sum AC<M>, sites passed, sites) are summed with ncclAllReduce -- the path's only collective --
 * and reported on stderr when BGT_B200_TOTALS=1.
 *
 * Anything else (-r/-B/-i/-n/-a/-t/-b/-u, several files, _mgs-masked genotypes, `**` filters, unordered records) goes to
 * ref_main_view, i.e. seam B.  The decision is taken before the first output byte; a device failure ends the run with
 * exit status 1 and a message -- it is never papered over by the CPU path.
 */
#include <fcntl.h>
#include <getopt.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <time.h>
#include "bgt.h"
#include "../include/bgt_b200.h"

int ref_main_view(int argc, char *argv[]);
b200_ctx_t *pbf_b200_ctx_dev(int dev);                          /* pbwt_shim.c */
const uint8_t *pbf_b200_image(const pbf_t *pb, size_t *len);
void pbf_b200_route_add(int slot, int64_t n);
void pbf_b200_route_report(void);

#define MAX_GPUS 16

static const uint8_t *map_file(const char *prefix, const char *ext, size_t *len)
{
	char *fn = (char*)malloc(strlen(prefix) + strlen(ext) + 1);
	struct stat sb;
	void *p;
	int fd;
	strcat(strcpy(fn, prefix), ext);
	fd = open(fn, O_RDONLY);
	free(fn);
	if (fd < 0) return 0;
	if (fstat(fd, &sb) != 0 || sb.st_size == 0) { close(fd); return 0; }
	p = mmap(0, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (p == MAP_FAILED) return 0;
	*len = (size_t)sb.st_size;
	return (const uint8_t*)p;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
#define TRACE(what) do { if (trace) { const double t_ = now_s(); fprintf(stderr, "[view_fast] %-28s %7.1f ms\n", what, 1e3 * (t_ - t_last)); t_last = t_; } } while (0)

static int run_reference(int argc, char *argv[])
{
	optind = 1;                                                  /* view.c runs its own getopt over the same argv */
	pbf_b200_route_add(1, 1);
	return ref_main_view(argc, argv);
}

/* what all shard workers share */
typedef struct {
	int n_gpus, dev[MAX_GPUS];
	const uint8_t *pbf, *bcf, *csi; size_t n_pbf, n_bcf, n_csi;
	bgtm_t *bm; bgt_file_t *file;
	const char *site_flt;
	const char **ctg; int n_ctg;
	unsigned vflags; int want_gt, trace;
	int64_t n_rows; int shift;
	/* phase 1 -> main: */
	pthread_barrier_t ready;
	/* ordered output */
	pthread_mutex_t lock; pthread_cond_t cond;
	int turn;                                                    /* shard whose text goes out next */
	int go;                                                      /* set by main after phase 1: 1 = write, -1 = abandon */
	htsFile *out;
} shared_t;

typedef struct {
	shared_t *sh; int t;
	pthread_t th;
	b200_ctx_t *ctx; int ctx_owned; b200_pbf_t *pb; b200_query_t *q; b200_sites_t *sites;
	int64_t row_beg, row_end, rec_beg, rec_end;
	int err, fallback;                                           /* phase-1 outcome */
	int64_t totals[5];                                           /* sum AN, sum AC, sum AC<M>, lines written, records scanned */
	int64_t n_lines;
} shard_t;

/* block-aligned row range of shard t of P (the same cut as bgt_b200/shard.py: ceil(blocks / P) blocks per shard) */
static void shard_rows(int64_t n, int shift, int P, int t, int64_t *beg, int64_t *end)
{
	const int64_t BS = 1LL << shift, nblk = (n + BS - 1) >> shift, per = (nblk + P - 1) / P;
	int64_t b = t * per * BS, e = (t + 1) * per * BS;
	if (b > n) b = n;
	if (e > n) e = n;
	*beg = b; *end = e;
}

static void shard_release(shard_t *s)
{
	if (s->sites) b200_sites_destroy(s->sites);
	if (s->q) b200_query_destroy(s->q);
	if (s->pb) b200_pbf_close(s->pb);
	s->sites = 0; s->q = 0; s->pb = 0;
	if (s->ctx_owned) b200_ctx_destroy(s->ctx);
	s->ctx = 0;
}

/* phase 1: everything that can fail or demand the reference path, before any output exists */
static void shard_prepare(shard_t *s)
{
	shared_t *sh = s->sh;
	bgtm_t *bm = sh->bm;
	int ferr = 0;
	double t_last = now_s();
	const int trace = sh->trace && s->t == 0;
	/* shard 0 shares the process-wide context of the seam; the others own theirs (one context per host thread, also when
	 * a device is listed more than once) */
	if (s->t == 0) s->ctx = pbf_b200_ctx_dev(sh->dev[0]);
	else if ((s->ctx = b200_ctx_create(sh->dev[s->t])) != 0) s->ctx_owned = 1;
	else fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror());
	if (s->ctx == 0) { s->err = 1; return; }
	TRACE("CUDA context");
	shard_rows(sh->n_rows, sh->shift, sh->n_gpus, s->t, &s->row_beg, &s->row_end);
	s->pb = b200_pbf_load_ex(s->ctx, sh->pbf, sh->n_pbf, s->row_beg, s->row_end, B200_LOAD_PREPARE_COUNT_SCAN);
	if (s->pb == 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); s->err = 1; return; }
	s->q = b200_query_create(s->ctx, s->pb, bm->bgt[0]->n_out, bm->bgt[0]->out, bm->group, bm->n_groups, sh->site_flt, &ferr);
	if (s->q == 0) {
		if (ferr) { s->fallback = 1; return; }                   /* kexpr took the filter, the device compiler did not: seam B evaluates it on the host */
		fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); s->err = 1; return;
	}
	if (b200_query_filter_needs_host(s->q)) { s->fallback = 1; return; }   /* `**`: host libm, seam B */
	TRACE("PBF load + query");
	s->sites = b200_sites_load(s->ctx, sh->bcf, sh->n_bcf, sh->csi, sh->n_csi, bcf_id2int(sh->file->h0, BCF_DT_ID, "_row"));
	if (s->sites == 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); s->err = 1; return; }
	TRACE("sites: inflate, index, parse");
	{ /* the records of this shard: those whose row lies in [row_beg, row_end) -- needs the records in row order unless one GPU formats the whole file at once */
		const int64_t n_rec = b200_sites_n(s->sites);
		s->rec_beg = 0; s->rec_end = n_rec;
		if (!b200_sites_rows_sorted(s->sites)) {
			if (sh->n_gpus > 1 || sh->want_gt) s->fallback = 1;
		} else if (sh->n_gpus > 1) {
			if (b200_sites_rec_range(s->sites, s->row_beg, s->row_end, &s->rec_beg, &s->rec_end) != 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); s->err = 1; }
		}
	}
}

/* phase 2: the shard's text, window by window, written when it is this shard's turn */
static void shard_emit(shard_t *s)
{
	shared_t *sh = s->sh;
	bgtm_t *bm = sh->bm;
	int64_t per = s->rec_end - s->rec_beg, beg;
	int turn_taken = 0;
	if (sh->want_gt) { per = (256LL << 20) / (4LL * bm->bgt[0]->n_out + 96); if (per < 1) per = 1; }   /* 4 bytes per sample and record */
	for (beg = s->rec_beg; beg < s->rec_end; beg += per) {
		const int64_t end = beg + per < s->rec_end ? beg + per : s->rec_end;
		const char *text = 0;
		int64_t n_lines = 0, tot[5], len;
		len = b200_view_text_ex(s->ctx, s->sites, s->pb, s->q, sh->vflags, beg, end, sh->ctg, sh->n_ctg, &text, &n_lines);
		if (len < 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); s->err = 1; break; }
		if (b200_last_totals(s->ctx, tot) == 0) { s->totals[0] += tot[0]; s->totals[1] += tot[1]; s->totals[2] += tot[2]; }
		s->totals[3] += n_lines; s->totals[4] += end - beg;
		if (!turn_taken) { /* the text buffer belongs to the sites handle until the next call: wait for the shards in front */
			pthread_mutex_lock(&sh->lock);
			while (sh->turn != s->t) pthread_cond_wait(&sh->cond, &sh->lock);
			pthread_mutex_unlock(&sh->lock);
			turn_taken = 1;
		}
		if (len > 0 && fwrite(text, 1, (size_t)len, (FILE*)sh->out->fp) != (size_t)len) { fprintf(stderr, "[E::%s] write failed\n", __func__); s->err = 1; break; }
	}
	pthread_mutex_lock(&sh->lock);
	while (sh->turn != s->t) pthread_cond_wait(&sh->cond, &sh->lock);
	++sh->turn;
	pthread_cond_broadcast(&sh->cond);
	pthread_mutex_unlock(&sh->lock);
}

static void *shard_main(void *arg)
{
	shard_t *s = (shard_t*)arg;
	shared_t *sh = s->sh;
	shard_prepare(s);
	pthread_barrier_wait(&sh->ready);                            /* main looks at every shard's outcome ... */
	pthread_mutex_lock(&sh->lock);
	while (sh->go == 0) pthread_cond_wait(&sh->cond, &sh->lock);  /* ... and decides */
	pthread_mutex_unlock(&sh->lock);
	if (sh->go > 0) shard_emit(s);
	return 0;
}

static int parse_devices(int dev[MAX_GPUS])
{
	const char *e = getenv("BGT_B200_DEVICES"), *one = getenv("BGT_B200_DEVICE");
	int n = 0;
	if (e == 0 || *e == 0) { dev[0] = one && *one ? atoi(one) : 0; return 1; }
	if (strcmp(e, "all") == 0) {
		n = b200_device_count();
		if (n > MAX_GPUS) n = MAX_GPUS;
		for (int i = 0; i < n; ++i) dev[i] = i;
		return n > 0 ? n : 1;
	}
	while (*e && n < MAX_GPUS) {
		char *end;
		long v = strtol(e, &end, 10);
		if (end == e) break;
		dev[n++] = (int)v;
		e = *end == ',' ? end + 1 : end;
	}
	if (n == 0) { dev[0] = 0; n = 1; }
	return n;
}

int main_view(int argc, char *argv[])
{
	int c, n_groups = 0, multi_flag = 0, other = 0, i, ret = 1;
	char *gexpr[BGT_MAX_GROUPS], *site_flt = 0;
	const char *off = getenv("BGT_B200_DISABLE"), *nofast = getenv("BGT_B200_NO_FASTVIEW");
	char **av = (char**)malloc((argc + 1) * sizeof(char*));
	memcpy(av, argv, (argc + 1) * sizeof(char*));               /* getopt permutes its argv: look at a copy */
	while ((c = getopt(argc, av, "ubs:r:l:CMGB:ef:g:a:i:n:SHt:d:")) >= 0) {   /* view.c:28 */
		if (c == 'C') multi_flag |= BGT_F_SET_AC;
		else if (c == 'G') multi_flag |= BGT_F_NO_GT;
		else if (c == 'f') site_flt = optarg;
		else if (c == 's' && n_groups < BGT_MAX_GROUPS) gexpr[n_groups++] = optarg;
		else if (c == 'l' || c == 'M') ;                         /* no effect on text output */
		else other = 1;
	}
	i = argc - optind;
	if (other || i != 1 || (off && *off == '1') || (nofast && *nofast == '1')) { free(av); optind = 1; return ref_main_view(argc, argv); }   /* seam B */
	{
		const char *prefix = av[optind];
		bgt_file_t *file;
		bgtm_t *bm;
		shared_t sh;
		shard_t sd[MAX_GPUS];
		int fallback = 0, failed = 0, t;
		const int trace = getenv("BGT_B200_TRACE") != 0;
		double t_last = now_s();
		memset(&sh, 0, sizeof(sh));
		memset(sd, 0, sizeof(sd));
		if (n_groups > 1) multi_flag |= BGT_F_SET_AC;            /* view.c:54 */
		if ((file = bgt_open(prefix)) == 0) {                    /* view.c:102-107 */
			fprintf(stderr, "[E::%s] failed to open BGT with prefix '%s'\n", __func__, prefix);
			free(av);
			return 1;
		}
		bm = bgtm_reader_init(1, &file);
		bgtm_set_flag(bm, multi_flag);
		if (site_flt && bgtm_set_flt_site(bm, site_flt) != 0) {  /* view.c:111-114 */
			fprintf(stderr, "[E::%s] failed to set frequency filters. Syntax error?\n", __func__);
			free(av);
			return 1;
		}
		for (i = 0; i < n_groups; ++i)
			if (bgtm_add_group(bm, gexpr[i]) < 0) {              /* view.c:134-139 */
				fprintf(stderr, "[E::%s] failed to add sample group '%s'.\n", __func__, gexpr[i]);
				free(av);
				return 1;
			}
		TRACE("bgt_open + reader setup");
		bgtm_prepare(bm);                                        /* generates the VCF header (view.c:140) */
		TRACE("bgtm_prepare");
		pbf_b200_route_add(0, 1);
		sh.bm = bm; sh.file = file; sh.site_flt = site_flt; sh.trace = trace;
		sh.pbf = pbf_b200_image(bm->bgt[0]->pb, &sh.n_pbf);
		sh.bcf = map_file(prefix, ".bcf", &sh.n_bcf);
		sh.csi = map_file(prefix, ".bcf.csi", &sh.n_csi);
		sh.want_gt = !(multi_flag & BGT_F_NO_GT);
		sh.vflags = ((multi_flag & BGT_F_SET_AC) ? B200_VIEW_COUNTS : 0) | (sh.want_gt ? B200_VIEW_GENOTYPES : 0);
		if (bm->bgt[0]->n_out <= 0 || sh.pbf == 0 || sh.bcf == 0) fallback = 1;
		if (sh.want_gt && bm->mgs)                               /* minimal-group-size masking of genotypes (bgt.c:294-307) stays the reference's */
			for (i = 0; i < bm->n_out; ++i) if (bm->mgs[i] > 1) fallback = 1;
		if (!fallback) {
			sh.n_rows = pbf_get_n(bm->bgt[0]->pb); sh.shift = pbf_get_shift(bm->bgt[0]->pb);
			sh.n_gpus = parse_devices(sh.dev);
			{ /* never more shards than checkpoint blocks */
				const int64_t nblk = (sh.n_rows + (1LL << sh.shift) - 1) >> sh.shift;
				if (sh.n_gpus > nblk) sh.n_gpus = nblk > 0 ? (int)nblk : 1;
			}
			sh.n_ctg = bm->h_out->n[BCF_DT_CTG];
			sh.ctg = (const char**)malloc((sh.n_ctg + 1) * sizeof(char*));
			for (i = 0; i < sh.n_ctg; ++i) sh.ctg[i] = bm->h_out->id[BCF_DT_CTG][i].key;
			pthread_mutex_init(&sh.lock, 0); pthread_cond_init(&sh.cond, 0);
			pthread_barrier_init(&sh.ready, 0, (unsigned)sh.n_gpus);
			for (t = 0; t < sh.n_gpus; ++t) { sd[t].sh = &sh; sd[t].t = t; }
			for (t = 1; t < sh.n_gpus; ++t) pthread_create(&sd[t].th, 0, shard_main, &sd[t]);
			shard_prepare(&sd[0]);                               /* shard 0 on this thread */
			pthread_barrier_wait(&sh.ready);
			for (t = 0; t < sh.n_gpus; ++t) { failed |= sd[t].err; fallback |= sd[t].fallback; }
			if (!failed && !fallback) {
				TRACE("all shards prepared");
				sh.out = hts_open("-", "w-1", 0);                /* view.c:142-147 */
				vcf_hdr_write(sh.out, bm->h_out);
			}
			pthread_mutex_lock(&sh.lock);
			sh.go = (failed || fallback) ? -1 : 1;
			pthread_cond_broadcast(&sh.cond);
			pthread_mutex_unlock(&sh.lock);
			if (sh.go > 0) shard_emit(&sd[0]);
			for (t = 1; t < sh.n_gpus; ++t) pthread_join(sd[t].th, 0);
			for (t = 0; t < sh.n_gpus; ++t) failed |= sd[t].err;
			if (sh.out) { hts_close(sh.out); TRACE("scan + text + write"); }
			pbf_b200_route_add(6, sh.n_gpus);
			if (!failed && !fallback) {
				const char *want_tot = getenv("BGT_B200_TOTALS");
				ret = 0;
				if (want_tot && *want_tot == '1') { /* whole-cohort totals: the path's one collective */
					int64_t tot[MAX_GPUS][5];
					b200_ctx_t *cx[MAX_GPUS];
					int dup = 0, u;
					for (t = 0; t < sh.n_gpus; ++t) { memcpy(tot[t], sd[t].totals, sizeof(tot[t])); cx[t] = sd[t].ctx; }
					for (t = 0; t < sh.n_gpus; ++t) for (u = 0; u < t; ++u) if (sh.dev[u] == sh.dev[t]) dup = 1;
					if (dup) { /* a device listed twice (testing the shard logic on one GPU): NCCL wants distinct devices */
						for (t = 1; t < sh.n_gpus; ++t) for (u = 0; u < 5; ++u) tot[0][u] += tot[t][u];
					} else if (sh.n_gpus > 1 && b200_allreduce_i64(cx, sh.n_gpus, &tot[0][0], 5) != 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); ret = 1; }
					if (ret == 0)
						fprintf(stderr, "[b200 totals] gpus=%d sites=%lld passed=%lld sum_AN=%lld sum_AC=%lld sum_AC2=%lld%s\n", sh.n_gpus, (long long)tot[0][4],
						        (long long)tot[0][3], (long long)tot[0][0], (long long)tot[0][1], (long long)tot[0][2],
						        sh.n_gpus > 1 ? (dup ? " (host sum: duplicate devices)" : " (ncclAllReduce)") : "");
					TRACE("totals all-reduce");
				}
			}
			/* (leaving through _exit once the output is flushed was measured: the kernel driver then reclaims the process's GPU
			 * state on its own slow path -- 1.9 s wall instead of 0.64 s; the orderly release below is the fast way out) */
			for (t = 0; t < sh.n_gpus; ++t) shard_release(&sd[t]);
			pthread_barrier_destroy(&sh.ready); pthread_mutex_destroy(&sh.lock); pthread_cond_destroy(&sh.cond);
			free(sh.ctg);
		}
		if (sh.bcf) munmap((void*)sh.bcf, sh.n_bcf);
		if (sh.csi) munmap((void*)sh.csi, sh.n_csi);
		bgtm_reader_destroy(bm);
		bgt_close(file);
		free(av);
		if (fallback && !failed) return run_reference(argc, argv);
		return failed ? 1 : ret;
	}
}
