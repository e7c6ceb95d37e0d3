/*
 * bgt_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * CPU restatement of the per-site AC/AN reduction and filter verdict of `bgt view`
 * (bgt.c:207-246, 333-345, 692-757, 850-857), driven over a memory-resident PBF.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "oracle.h"

/* bgt.c:735-757: histogram of the 2-bit code a1<<1|a0 over the tracked haplotypes, per group when
 * there is more than one; AN = c0+c1+c3, AC = {c1, c3}.
 * Deviation, stated: with ONE group the reference leaves gan[0]/gac[0] uninitialised although
 * bgtm_assign_expr (bgt.c:706-709) binds AN1/AC1 from them; here they mirror AN/AC. */
void orc_cal_info(int n_out, const uint8_t *a0, const uint8_t *a1, const uint32_t *group, int n_groups, orc_info_t *ss)
{
	int32_t tot[4] = {0, 0, 0, 0};
	int i, j;
	memset(ss, 0, sizeof(*ss));
	ss->n_groups = n_groups;
	if (n_groups > 1) {
		int32_t gc[ORC_MAX_GROUPS][4];
		memset(gc, 0, sizeof(gc));
		for (i = 0; i < 2 * n_out; ++i)
			++gc[group[i >> 1] - 1][a1[i] << 1 | a0[i]];
		for (i = 0; i < n_groups; ++i) {
			ss->gan[i] = gc[i][0] + gc[i][1] + gc[i][3];
			ss->gac[i][0] = gc[i][1];
			ss->gac[i][1] = gc[i][3];
			for (j = 0; j < 4; ++j) tot[j] += gc[i][j];
		}
	} else {
		for (i = 0; i < 2 * n_out; ++i) ++tot[a1[i] << 1 | a0[i]];
	}
	ss->an = tot[0] + tot[1] + tot[3];
	ss->ac[0] = tot[1]; ss->ac[1] = tot[3];
	if (n_groups <= 1) { ss->gan[0] = ss->an; ss->gac[0][0] = ss->ac[0]; ss->gac[0][1] = ss->ac[1]; }
}

/* bgt.c:692-698: "AN1".."AN9", "AN10".. */
static void group_key(char key[5], char nc, int g)
{
	key[0] = 'A'; key[1] = nc;
	if (g < 9) key[2] = (char)('0' + g + 1), key[3] = 0;
	else key[2] = (char)('0' + (g + 1) / 10), key[3] = (char)('0' + (g + 1) % 10), key[4] = 0;
}

/* bgt.c:700-719: bind AN, AC (first ALT only), AN#, AC#; pass iff no evaluation error and int result != 0 */
int orc_pass_site(const orc_info_t *ss, orc_expr_t *flt)
{
	int i, err, vt;
	int64_t iv; double rv;
	char key[5];
	if (flt == 0) return 1;
	orc_expr_set_int(flt, "AN", ss->an);
	orc_expr_set_int(flt, "AC", ss->ac[0]);
	for (i = 0; i < ss->n_groups; ++i) {
		group_key(key, 'N', i); orc_expr_set_int(flt, key, ss->gan[i]);
		group_key(key, 'C', i); orc_expr_set_int(flt, key, ss->gac[i][0]);
	}
	err = orc_expr_eval(flt, &iv, &rv, &vt);
	return err ? 0 : (iv != 0);
}

int64_t orc_scan(orc_pbf_t *p, int64_t row_beg, int64_t n_rows, int n_out, const int32_t *out_samples,
                 const uint32_t *group, int n_groups, const char *flt,
                 int32_t *counts, uint8_t *pass, uint8_t *hap0, uint8_t *hap1)
{
	int m = orc_pbf_m(p), n_track = 2 * n_out, i, err = 0, stride = 3 + 3 * n_groups;
	int *cols = (int*)malloc((size_t)(n_track ? n_track : 1) * sizeof(int));
	orc_expr_t *ke = 0;
	int64_t k, done = 0;
	if (orc_pbf_g(p) != 2 || n_groups < 1 || n_groups > ORC_MAX_GROUPS) { free(cols); return -1; }
	if (flt) { ke = orc_expr_parse(flt, &err); if (err || !ke) { free(cols); return -2; } }
	for (i = 0; i < n_out; ++i) /* bgt.c:239-242 */
		cols[2*i] = out_samples[i] << 1, cols[2*i+1] = out_samples[i] << 1 | 1;
	orc_pbf_subset(p, n_track, cols); /* n_track >= m  =>  full decode (pbwt.c:377) */
	if (orc_pbf_seek(p, row_beg) < 0 && row_beg != 0) { free(cols); orc_expr_free(ke); return -3; }
	(void)m;
	for (k = 0; k < n_rows; ++k) {
		const uint8_t **a = orc_pbf_read(p);
		orc_info_t ss;
		int32_t *c;
		if (!a) break;
		orc_cal_info(n_out, a[0], a[1], group, n_groups, &ss);
		c = counts + k * stride;
		c[0] = ss.an; c[1] = ss.ac[0]; c[2] = ss.ac[1];
		for (i = 0; i < n_groups; ++i)
			c[3+3*i] = ss.gan[i], c[4+3*i] = ss.gac[i][0], c[5+3*i] = ss.gac[i][1];
		if (pass) pass[k] = (uint8_t)orc_pass_site(&ss, ke);
		if (hap0) memcpy(hap0 + k * (int64_t)n_track, a[0], n_track);
		if (hap1) memcpy(hap1 + k * (int64_t)n_track, a[1], n_track);
		++done;
	}
	free(cols); orc_expr_free(ke);
	return done;
}
