/*
 * view_fast.c -- `bgt view -G` with the whole per-site pipeline on the device.
 *
 * The reference's view.c is compiled unchanged with -Dmain_view=ref_main_view; this main_view looks at the options
 * first.  For the VCF scan of one BGT -- `view [-G] [-C] [-f EXPR] [-s EXPR ...] prefix`, BASELINE configs 2-5 -- set-up
 * and the header are the reference's own calls (bgt_open, bgtm_reader_init, bgtm_set_flag,
 * bgtm_set_flt_site, bgtm_add_group, bgtm_prepare, vcf_hdr_write: view.c:99-147), and the record loop of
 * view.c:150-155 (bgtm_read + vcf_write1 per site on the host thread) is replaced by ONE call: the .bcf/.csi are
 * inflated, indexed and parsed on the GPU (b200_sites_load), the .pbf is scanned (b200_view_text -> b200_scan) and the
 * VCF lines of the passing sites come back as text, window by window when genotype columns are printed.  Anything else
 * (-r/-B/-i/-n/-a/-t/-b/-u, several files, _mgs-masked genotypes, filters the device compiler rejects) goes to
 * ref_main_view, i.e. seam B.
 */
#include <fcntl.h>
#include <getopt.h>
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
b200_ctx_t *pbf_b200_ctx(void);                                  /* pbwt_shim.c */
const uint8_t *pbf_b200_image(const pbf_t *pb, size_t *len);

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
	return ref_main_view(argc, argv);
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
	if (other || i != 1 || (off && *off == '1') || (nofast && *nofast == '1')) { free(av); return run_reference(argc, argv); }
	{
		const char *prefix = av[optind];
		bgt_file_t *file;
		bgtm_t *bm;
		htsFile *out;
		b200_ctx_t *ctx;
		b200_sites_t *sites = 0;
		b200_pbf_t *pb = 0;
		b200_query_t *q = 0;
		const uint8_t *bcf = 0, *csi = 0, *pbf;
		size_t n_bcf = 0, n_csi = 0, n_pbf = 0;
		const char **ctg = 0, *text = 0;
		int64_t len, n_lines = 0;
		int err = 0, fallback = 0;
		const int trace = getenv("BGT_B200_TRACE") != 0;
		double t_last = now_s();
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
		ctx = pbf_b200_ctx();
		TRACE("CUDA context");
		pbf = pbf_b200_image(bm->bgt[0]->pb, &n_pbf);
		bcf = map_file(prefix, ".bcf", &n_bcf);
		csi = map_file(prefix, ".bcf.csi", &n_csi);
		if (bm->bgt[0]->n_out <= 0 || pbf == 0 || bcf == 0) fallback = 1;
		if (!(multi_flag & BGT_F_NO_GT) && bm->mgs)                /* minimal-group-size masking of genotypes (bgt.c:294-307) stays the reference's */
			for (i = 0; i < bm->n_out; ++i) if (bm->mgs[i] > 1) fallback = 1;
		if (!fallback) {
			pb = b200_pbf_load_ex(ctx, pbf, n_pbf, 0, -1, B200_LOAD_PREPARE_COUNT_SCAN);
			if (pb == 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); exit(1); }
			q = b200_query_create(ctx, pb, bm->bgt[0]->n_out, bm->bgt[0]->out, bm->group, bm->n_groups, site_flt, &err);
			if (q == 0) fallback = 1;                            /* kexpr took the filter but the device compiler did not */
			TRACE("PBF load + query");
		}
		if (!fallback) {
			sites = b200_sites_load(ctx, bcf, n_bcf, csi, n_csi, bcf_id2int(file->h0, BCF_DT_ID, "_row"));
			if (sites == 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); exit(1); }
			TRACE("sites: inflate, index, parse");
			ctg = (const char**)malloc((bm->h_out->n[BCF_DT_CTG] + 1) * sizeof(char*));
			for (i = 0; i < bm->h_out->n[BCF_DT_CTG]; ++i) ctg[i] = bm->h_out->id[BCF_DT_CTG][i].key;
			{
				/* records in windows that bound the text per call: 4 bytes per sample and record when genotypes are printed */
				const int want_gt = !(multi_flag & BGT_F_NO_GT);
				const unsigned vflags = ((multi_flag & BGT_F_SET_AC) ? B200_VIEW_COUNTS : 0) | (want_gt ? B200_VIEW_GENOTYPES : 0);
				const int64_t n_rec = b200_sites_n(sites);
				int64_t per = n_rec, beg;
				if (want_gt) { per = (256LL << 20) / (4LL * bm->bgt[0]->n_out + 96); if (per < 1) per = 1; }
				out = 0;
				for (beg = 0; beg < n_rec || beg == 0; beg += per) {
					const int64_t end = beg + per < n_rec ? beg + per : n_rec;
					len = b200_view_text_ex(ctx, sites, pb, q, vflags, beg, end, ctg, bm->h_out->n[BCF_DT_CTG], &text, &n_lines);
					if (len < 0) {
						if (beg == 0 && (strstr(b200_strerror(), "host libm") || strstr(b200_strerror(), "row order"))) { fallback = 1; break; }   /* `**` filters, unordered records: seam B */
						fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror());
						exit(1);
					}
					if (out == 0) {
						TRACE("first window: scan + text");
						out = hts_open("-", "w-1", 0);               /* view.c:142-147 */
						vcf_hdr_write(out, bm->h_out);
					}
					if (len > 0 && fwrite(text, 1, (size_t)len, (FILE*)out->fp) != (size_t)len) { fprintf(stderr, "[E::%s] write failed\n", __func__); exit(1); }
					if (n_rec == 0) break;
				}
				if (out) { hts_close(out); TRACE("remaining windows + write"); ret = 0; }
			}
		}
		free(ctg);
		if (sites) b200_sites_destroy(sites);
		if (q) b200_query_destroy(q);
		if (pb) b200_pbf_close(pb);
		if (bcf) munmap((void*)bcf, n_bcf);
		if (csi) munmap((void*)csi, n_csi);
		bgtm_reader_destroy(bm);
		bgt_close(file);
		free(av);
		if (fallback) return run_reference(argc, argv);
		return ret;
	}
}
