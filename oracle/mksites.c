/*
 * mksites.c -- TEST INFRASTRUCTURE ONLY.
 * Writes the SITE side of a synthetic BGT database for an existing <prefix>.pbf:
 *   <prefix>.bcf (site-only BCF2 with INFO/_row), <prefix>.bcf.csi (+RNI) and <prefix>.spl,
 * entirely through the UNMODIFIED reference library (oracle/_ref/libbgt.a: vcf_parse1, bcf_append_info_ints,
 * vcf_write1, bcf_index_build -- the calls import.c:49-117 makes), so that the reference `bgt view` can be run
 * on cohorts whose genotype side came from the B200 generator.  Layout per SURVEY 8d: one contig, POS = 1000 +
 * 10*row, single-base REF/ALT, a second ALT "<M>" on rows whose plane 1 is not empty, samples S%07d with
 * grp:Z:A (even index) / grp:Z:B (odd) and idx:i:<index>.
 *
 * usage: mksites <prefix>      (reads <prefix>.pbf for m, n and the per-row plane-1 content)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "vcf.h"
#include "kstring.h"

int main(int argc, char *argv[])
{
	char *fn;
	FILE *fp;
	int32_t hdr[3], m, g;
	int64_t n = 0, k;
	uint8_t t, *buf;
	htsFile *out;
	bcf_hdr_t *h;
	bcf1_t *b;
	kstring_t s = {0, 0, 0};
	const char *bases = "ACGT";
	if (argc < 2) { fprintf(stderr, "usage: mksites <prefix>\n"); return 1; }
	fn = (char*)malloc(strlen(argv[1]) + 16);
	sprintf(fn, "%s.pbf", argv[1]);
	if ((fp = fopen(fn, "rb")) == 0) { fprintf(stderr, "cannot open %s\n", fn); return 1; }
	{ char magic[4]; if (fread(magic, 1, 4, fp) != 4) return 1; } /* checked by the readers */
	if (fread(hdr, 4, 3, fp) != 3) return 1;
	m = hdr[0]; g = hdr[1];
	buf = (uint8_t*)malloc((size_t)m + 8);

	h = bcf_hdr_init();
	kputs("##fileformat=VCFv4.1\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n", &s);
	kputs("##contig=<ID=11,length=135006516>\n", &s);
	kputs("##INFO=<ID=_row,Number=1,Type=Integer,Description=\"row number\">\n", &s);
	kputs("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n", &s);
	h->text = s.s; h->l_text = s.l + 1; s.s = 0; s.l = s.m = 0;
	bcf_hdr_parse(h);
	sprintf(fn, "%s.bcf", argv[1]);
	out = hts_open(fn, "wb", 0);
	vcf_hdr_write(out, h);
	b = bcf_init1();
	for (k = 0;; ++k) {
		int32_t l, val = (int32_t)k, p, multi = 0;
		if (fread(&t, 1, 1, fp) != 1) break;
		if (t == 'S') { fseek(fp, (long)g * 4 * m, SEEK_CUR); if (fread(&t, 1, 1, fp) != 1) break; }
		if (t != 'B') break;
		for (p = 0; p < g; ++p) {
			int i;
			if (fread(&l, 4, 1, fp) != 1 || fread(buf, 1, l, fp) != (size_t)l) return 1;
			if (p == 1) for (i = 0; i < l; ++i) if ((buf[i] & 1) && (buf[i] >> 1)) multi = 1;
		}
		s.l = 0;
		ksprintf(&s, "11\t%ld\t.\t%c\t%c%s\t0\t.\t.", (long)(1000 + 10 * k), bases[k & 3], bases[(k + 1) & 3], multi ? ",<M>" : "");
		vcf_parse1(&s, h, b);
		bcf_append_info_ints(h, b, "_row", 1, &val);
		vcf_write1(out, h, b);
		++n;
	}
	hts_close(out);
	fclose(fp);
	bcf_index_build(fn, 14);
	sprintf(fn, "%s.spl", argv[1]);
	fp = fopen(fn, "wb");
	for (k = 0; k < m / 2; ++k) fprintf(fp, "S%07ld\tgrp:Z:%c\tidx:i:%ld\n", (long)k, (k & 1) ? 'B' : 'A', (long)k);
	fclose(fp);
	fprintf(stderr, "[mksites] %ld sites, %d samples\n", (long)n, m / 2);
	return 0;
}
