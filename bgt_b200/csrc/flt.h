// flt.h -- site-filter byte-code shared by the host compiler (flt.cpp) and the device evaluator.
//
// `bgt view -f EXPR` (view.c:42, bgt.c:444-455) parses EXPR with kexpr and evaluates it once per site with
// AN, AC, AN<i>, AC<i> bound to the site's counts (bgt.c:700-719).  The expression is compiled ON THE HOST,
// once, into a postfix program; the per-site evaluation -- the part that scales with the number of sites --
// runs on the device inside the finalize kernel.  Value semantics follow kexpr.c:78-160 exactly: every value
// carries an int64 and a double side by side plus a type tag, `/` is always real, comparisons are real when
// either side is real, and any unbound variable or undefined function fails the site (kexpr.c:371-374).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define FLT_HD __host__ __device__ __forceinline__
#else
#define FLT_HD inline
#endif

#define FLT_MAX_CODE   96
#define FLT_MAX_STACK  24
#define FLT_MAX_STR    8

// operator numbering = kexpr.c:14-38
enum { FO_NULL, FO_POS, FO_NEG, FO_BNOT, FO_LNOT, FO_POW, FO_MUL, FO_DIV, FO_IDIV, FO_MOD, FO_ADD, FO_SUB, FO_LSH,
       FO_RSH, FO_LT, FO_LE, FO_GT, FO_GE, FO_EQ, FO_NE, FO_BAND, FO_BXOR, FO_BOR, FO_LAND, FO_LOR, FO_ABS };
enum { FK_CONST = 1, FK_VAR = 2, FK_OP1 = 3, FK_OP2 = 4, FK_DROP = 5 };   // instruction kinds
enum { FV_REAL = 1, FV_INT = 2, FV_STR = 3 };                             // kexpr.h:23-25

struct flt_ins_t {
	uint8_t kind, op, vtype, sid;   // sid: string-constant id for FV_STR constants
	int32_t arg;                    // FK_VAR: slot in the per-site count vector; FK_DROP: stack entries to drop
	int64_t i;
	double r;
};

struct flt_prog_t {
	int32_t n;              // 0 = no filter (every site passes)
	int32_t always_fail;    // unbound variable / undefined function somewhere: kexpr reports an error at every site
	int32_t needs_host;     // contains `**` (libm pow): evaluated with the host's libm so results match the reference bit for bit
	int32_t n_str;
	int8_t  scmp[FLT_MAX_STR][FLT_MAX_STR]; // sign of strcmp between string constants
	flt_ins_t code[FLT_MAX_CODE];
};

struct flt_val_t { int64_t i; double r; int32_t vt, sid; };

// (int64_t)(r + .5) as the reference's x86-64 build computes it: cvttsd2si yields INT64_MIN for NaN and for
// values outside [-2^63, 2^63).
FLT_HD int64_t flt_r2i(double r)
{
	double x = r + .5;
	if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
	return (int64_t)x;
}

FLT_HD void flt_apply2(int op, flt_val_t *p, const flt_val_t *q, const flt_prog_t *P, int *fault)
{
	const bool real = (p->vt == FV_REAL || q->vt == FV_REAL);
	switch (op) {
	case FO_MUL: p->i = (int64_t)((uint64_t)p->i * (uint64_t)q->i); p->r *= q->r; p->vt = real ? FV_REAL : FV_INT; break;
	case FO_ADD: p->i = (int64_t)((uint64_t)p->i + (uint64_t)q->i); p->r += q->r; p->vt = real ? FV_REAL : FV_INT; break;
	case FO_SUB: p->i = (int64_t)((uint64_t)p->i - (uint64_t)q->i); p->r -= q->r; p->vt = real ? FV_REAL : FV_INT; break;
	case FO_DIV: p->r /= q->r; p->i = flt_r2i(p->r); p->vt = FV_REAL; break;
	case FO_POW: p->r = pow(p->r, q->r); p->i = flt_r2i(p->r); p->vt = real ? FV_REAL : FV_INT; break;
	case FO_IDIV: case FO_MOD:
		// the reference divides by q->i unguarded (SIGFPE on 0 or INT64_MIN/-1); such a site cannot be output by
		// the reference either, so it is failed here.
		if (q->i == 0 || (p->i == INT64_MIN && q->i == -1)) { *fault = 1; p->i = 0; }
		else p->i = (op == FO_IDIV) ? p->i / q->i : p->i % q->i;
		p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_LSH: p->i = (int64_t)((uint64_t)p->i << (q->i & 63)); p->r = (double)p->i; p->vt = FV_INT; break; // x86 shl masks the count
	case FO_RSH: p->i = p->i >> (q->i & 63); p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_BAND: p->i &= q->i; p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_BXOR: p->i ^= q->i; p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_BOR: p->i |= q->i; p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_LAND: p->i = (p->i && q->i); p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_LOR: p->i = (p->i || q->i); p->r = (double)p->i; p->vt = FV_INT; break;
	default: { // FO_LT..FO_NE, kexpr.c:78-92
		int c;
		if (p->vt == FV_STR && q->vt == FV_STR) {
			const int d = P->scmp[p->sid][q->sid];
			c = op == FO_LT ? d < 0 : op == FO_LE ? d <= 0 : op == FO_GT ? d > 0 : op == FO_GE ? d >= 0 : op == FO_EQ ? d == 0 : d != 0;
		} else if (real) {
			const double a = p->r, b = q->r;
			c = op == FO_LT ? a < b : op == FO_LE ? a <= b : op == FO_GT ? a > b : op == FO_GE ? a >= b : op == FO_EQ ? a == b : a != b;
		} else {
			const int64_t a = p->i, b = q->i;
			c = op == FO_LT ? a < b : op == FO_LE ? a <= b : op == FO_GT ? a > b : op == FO_GE ? a >= b : op == FO_EQ ? a == b : a != b;
		}
		p->i = c; p->r = (double)c; p->vt = FV_INT;
	} }
}

FLT_HD void flt_apply1(int op, flt_val_t *p)
{
	switch (op) {
	case FO_POS: break;
	case FO_NEG: p->i = (int64_t)(0 - (uint64_t)p->i); p->r = -p->r; break;   // type tag untouched, kexpr.c:153
	case FO_BNOT: p->i = ~p->i; p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_LNOT: p->i = !p->i; p->r = (double)p->i; p->vt = FV_INT; break;
	case FO_ABS: // kexpr.c:155: integer abs() goes through C's int abs(int)
		if (p->vt == FV_INT) { int32_t t = (int32_t)p->i; t = (t < 0 && t != INT32_MIN) ? -t : t; p->i = t; p->r = (double)p->i; }
		else { p->r = fabs(p->r); p->i = flt_r2i(p->r); }
		break;
	}
}

// vars: the site's count vector [AN, AC, AC<M>, AN1, AC1, AC1<M>, ...] (stride 3+3G).  Returns 1 = site passes.
FLT_HD int flt_eval(const flt_prog_t *P, const int32_t *vars)
{
	if (P->n == 0) return 1;
	if (P->always_fail) return 0;
	flt_val_t st[FLT_MAX_STACK];
	int top = 0, fault = 0;
	for (int k = 0; k < P->n; ++k) {
		const flt_ins_t &e = P->code[k];
		if (e.kind == FK_CONST) { st[top].i = e.i; st[top].r = e.r; st[top].vt = e.vtype; st[top].sid = e.sid; ++top; }
		else if (e.kind == FK_VAR) { const int32_t v = vars[e.arg]; st[top].i = v; st[top].r = (double)v; st[top].vt = FV_INT; st[top].sid = 0; ++top; }
		else if (e.kind == FK_OP2) { --top; flt_apply2(e.op, &st[top-1], &st[top], P, &fault); }
		else if (e.kind == FK_OP1) flt_apply1(e.op, &st[top-1]);
		else top -= e.arg;
	}
	return fault ? 0 : (st[0].i != 0);
}

// host side (flt.cpp): returns the kexpr-style parse error mask (0 = ok)
int flt_compile(const char *expr, int n_groups, flt_prog_t *prog);
