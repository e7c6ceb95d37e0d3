/*
 * kexpr_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * CPU restatement of the expression semantics `bgt view -f` relies on (kexpr.c), as far as a site
 * filter can reach them: integer/real/string literals, named variables, the 24 operators with the
 * reference's precedence and associativity, function-call syntax (only abs() is ever defined for a
 * site filter -- bgt.c:444-455 never installs the math functions), the dual (int64,double) value
 * track, and the evaluation error mask.  Pinned against oracle/_ref/libbgtref.so (ke_parse/ke_eval).
 */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <math.h>
#include "oracle.h"

/* kexpr.h:10-20 */
enum { E_UNQU = 1, E_UNLP = 2, E_UNRP = 4, E_UNOP = 8, E_FUNC = 0x10, E_ARG = 0x20, E_NUM = 0x40,
       E_UNFUNC = 0x40, E_UNVAR = 0x80 };
enum { V_REAL = 1, V_INT = 2, V_STR = 3 };          /* kexpr.h:23-25 */
enum { T_NULL, T_VAL, T_OP, T_FUNC };               /* kexpr.c:40-43 */
/* kexpr.c:14-38, same numbering */
enum { O_NULL, O_POS, O_NEG, O_BNOT, O_LNOT, O_POW, O_MUL, O_DIV, O_IDIV, O_MOD, O_ADD, O_SUB, O_LSH, O_RSH,
       O_LT, O_LE, O_GT, O_GE, O_EQ, O_NE, O_BAND, O_BXOR, O_BOR, O_LAND, O_LOR };

/* kexpr.c:62-76: precedence level (low = binds tighter) and right-associativity flag */
static const int prec[25]  = {0, 1,1,1,1, 2, 3,3,3,3, 4,4, 5,5, 6,6,6,6, 7,7, 8, 9, 10, 11, 12};
static const int rassoc[25] = {0, 1,1,1,1, 1, 0,0,0,0, 0,0, 0,0, 0,0,0,0, 0,0, 0, 0, 0, 0, 0};

typedef struct {
	int ttype, vtype, op, n_args, assigned, defined; /* defined: operator/function has an implementation */
	int is_abs, lparen;
	char *name, *s;
	double r;
	int64_t i;
} tok_t;

struct orc_expr_s { int n; tok_t *e; };

static char *dupn(const char *s, int n) { char *d = (char*)calloc(n + 1, 1); memcpy(d, s, n); return d; }

/* kexpr.c:163-244 */
static tok_t read_token(char *p, char **end, int *err, int last_is_val)
{
	tok_t e;
	char *q = p;
	memset(&e, 0, sizeof(e));
	if (isalpha((unsigned char)*p) || *p == '_') {
		while (*p && (*p == '_' || isalnum((unsigned char)*p))) ++p;
		if (*p == '(') e.ttype = T_FUNC, e.n_args = 1;
		else e.ttype = T_VAL, e.vtype = V_REAL;
		e.name = dupn(q, (int)(p - q));
		*end = p;
	} else if (isdigit((unsigned char)*p) || *p == '.') {
		char *pd, *pl;
		double y = strtod(q, &pd);
		long x = strtol(q, &pl, 0);
		e.ttype = T_VAL;
		if (q == pd && q == pl) *err |= E_NUM;
		else if (pd > pl) e.vtype = V_REAL, e.i = (int64_t)(y + .5), e.r = y, *end = pd;
		else e.vtype = V_INT, e.i = x, e.r = y, *end = pl;
	} else if (*p == '"' || *p == '\'') {
		int c = *p;
		for (++p; *p && *p != c; ++p) if (*p == '\\') ++p;
		if (*p == c) e.ttype = T_VAL, e.vtype = V_STR, e.s = dupn(q + 1, (int)(p - q - 1)), *end = p + 1;
		else *err |= E_UNQU, *end = p;
	} else {
		static const struct { const char *s; int op, n; } tbl[] = { /* two-character operators first */
			{"**", O_POW, 2}, {"//", O_IDIV, 2}, {"==", O_EQ, 2}, {"!=", O_NE, 2}, {"<>", O_NE, 2}, {">=", O_GE, 2},
			{"<=", O_LE, 2}, {">>", O_RSH, 2}, {"<<", O_LSH, 2}, {"||", O_LOR, 2}, {"&&", O_LAND, 2},
			{"*", O_MUL, 2}, {"/", O_DIV, 2}, {"%", O_MOD, 2}, {">", O_GT, 2}, {"<", O_LT, 2}, {"|", O_BOR, 2},
			{"&", O_BAND, 2}, {"^", O_BXOR, 2}, {"~", O_BNOT, 1}, {"!", O_LNOT, 1}, {0, 0, 0} };
		int k;
		e.ttype = T_OP; e.defined = 1;
		if (*p == '+') e.op = last_is_val ? O_ADD : O_POS, e.n_args = last_is_val ? 2 : 1, *end = q + 1;
		else if (*p == '-') e.op = last_is_val ? O_SUB : O_NEG, e.n_args = last_is_val ? 2 : 1, *end = q + 1;
		else {
			for (k = 0; tbl[k].s; ++k)
				if (strncmp(p, tbl[k].s, strlen(tbl[k].s)) == 0) break;
			if (tbl[k].s) e.op = tbl[k].op, e.n_args = tbl[k].n, *end = q + strlen(tbl[k].s);
			else e.ttype = T_NULL, *err |= E_UNOP;
		}
	}
	return e;
}

static tok_t *push(tok_t **a, int *n, int *m)
{
	if (*n == *m) {
		int old = *m;
		*m = *m ? *m * 2 : 8;
		*a = (tok_t*)realloc(*a, (size_t)*m * sizeof(tok_t));
		memset(*a + old, 0, (size_t)(*m - old) * sizeof(tok_t));
	}
	return &(*a)[(*n)++];
}

/* kexpr.c:257-352: shunting-yard to RPN after squeezing out white space */
orc_expr_t *orc_expr_parse(const char *str, int *err)
{
	char *s = strdup(str), *p, *q;
	tok_t *out = 0, *ops = 0;
	int n_out = 0, m_out = 0, n_op = 0, m_op = 0, last_is_val = 0, i;
	orc_expr_t *ke;
	*err = 0;
	for (p = q = s; *p; ++p) if (!isspace((unsigned char)*p)) *q++ = *p;
	*q = 0;
	p = s;
	while (*p) {
		if (*p == '(') {
			tok_t *t = push(&ops, &n_op, &m_op);
			t->lparen = 1; t->ttype = T_NULL; ++p;
		} else if (*p == ')') {
			while (n_op > 0 && !ops[n_op-1].lparen) *push(&out, &n_out, &m_out) = ops[--n_op];
			if (n_op == 0) { *err |= E_UNRP; break; }
			--n_op;
			if (n_op > 0 && ops[n_op-1].ttype == T_FUNC) {
				tok_t *u = push(&out, &n_out, &m_out);
				*u = ops[--n_op];
				if (u->n_args == 1 && strcmp(u->name, "abs") == 0) u->is_abs = u->defined = 1; /* kexpr.c:286 */
			}
			++p;
		} else if (*p == ',') {
			while (n_op > 0 && !ops[n_op-1].lparen) *push(&out, &n_out, &m_out) = ops[--n_op];
			if (n_op < 2 || ops[n_op-2].ttype != T_FUNC) { *err |= E_FUNC; break; }
			++ops[n_op-2].n_args; ++p;
		} else {
			tok_t v = read_token(p, &p, err, last_is_val);
			if (*err) { free(v.name); free(v.s); break; }
			if (v.ttype == T_VAL) *push(&out, &n_out, &m_out) = v, last_is_val = 1;
			else if (v.ttype == T_FUNC) *push(&ops, &n_op, &m_op) = v, last_is_val = 0;
			else if (v.ttype == T_OP) {
				while (n_op > 0 && ops[n_op-1].ttype == T_OP) { /* kexpr.c:319-325 */
					int top = prec[ops[n_op-1].op];
					if ((rassoc[v.op] && prec[v.op] <= top) || (!rassoc[v.op] && prec[v.op] < top)) break;
					*push(&out, &n_out, &m_out) = ops[--n_op];
				}
				*push(&ops, &n_op, &m_op) = v; last_is_val = 0;
			}
		}
	}
	if (*err == 0) {
		while (n_op > 0 && !ops[n_op-1].lparen) *push(&out, &n_out, &m_out) = ops[--n_op];
		if (n_op > 0) *err |= E_UNLP;
	}
	if (*err == 0) { /* kexpr.c:336-343 */
		int n = 0;
		for (i = 0; i < n_out; ++i) n += out[i].ttype == T_VAL ? 1 : -(out[i].n_args - 1);
		if (n != 1) *err |= E_ARG;
	}
	free(s);
	if (*err) {
		for (i = 0; i < n_out; ++i) free(out[i].name), free(out[i].s);
		for (i = 0; i < n_op; ++i) free(ops[i].name), free(ops[i].s);
		free(out); free(ops);
		return 0;
	}
	free(ops);
	ke = (orc_expr_t*)calloc(1, sizeof(*ke));
	ke->n = n_out; ke->e = out;
	return ke;
}

void orc_expr_free(orc_expr_t *e)
{
	int i;
	if (!e) return;
	for (i = 0; i < e->n; ++i) free(e->e[i].name), free(e->e[i].s);
	free(e->e); free(e);
}

/* kexpr.c:432-442 */
int orc_expr_set_int(orc_expr_t *e, const char *name, int64_t v)
{
	int i, n = 0;
	for (i = 0; i < e->n; ++i) {
		tok_t *t = &e->e[i];
		if (t->ttype == T_VAL && t->name && strcmp(t->name, name) == 0)
			t->i = v, t->r = (double)v, t->vtype = V_INT, t->assigned = 1, ++n;
	}
	return n;
}

static int64_t r2i(double r) { return (int64_t)(r + .5); } /* the reference's rounding, x86 cvttsd2si */

/* kexpr.c:78-153: the operator bodies */
static void apply(int op, tok_t *p, const tok_t *q)
{
	int real = q && (p->vtype == V_REAL || q->vtype == V_REAL);
	switch (op) {
	case O_POS: break;
	case O_NEG: p->i = -p->i; p->r = -p->r; break;
	case O_BNOT: p->i = ~p->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_LNOT: p->i = !p->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_POW: p->r = pow(p->r, q->r); p->i = r2i(p->r); p->vtype = real ? V_REAL : V_INT; break;
	case O_MUL: p->i *= q->i; p->r *= q->r; p->vtype = real ? V_REAL : V_INT; break;
	case O_ADD: p->i += q->i; p->r += q->r; p->vtype = real ? V_REAL : V_INT; break;
	case O_SUB: p->i -= q->i; p->r -= q->r; p->vtype = real ? V_REAL : V_INT; break;
	case O_DIV: p->r /= q->r; p->i = r2i(p->r); p->vtype = V_REAL; break;
	case O_IDIV: p->i /= q->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_MOD: p->i %= q->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_LSH: p->i <<= q->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_RSH: p->i >>= q->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_BAND: p->i &= q->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_BXOR: p->i ^= q->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_BOR: p->i |= q->i; p->r = (double)p->i; p->vtype = V_INT; break;
	case O_LAND: p->i = (p->i && q->i); p->r = (double)p->i; p->vtype = V_INT; break;
	case O_LOR: p->i = (p->i || q->i); p->r = (double)p->i; p->vtype = V_INT; break;
	default: { /* comparisons, kexpr.c:78-92 */
		int c;
		if (p->vtype == V_STR && q->vtype == V_STR) {
			int d = strcmp(p->s, q->s);
			c = op == O_LT ? d < 0 : op == O_LE ? d <= 0 : op == O_GT ? d > 0 : op == O_GE ? d >= 0 : op == O_EQ ? d == 0 : d != 0;
		} else if (real) {
			double a = p->r, b = q->r;
			c = op == O_LT ? a < b : op == O_LE ? a <= b : op == O_GT ? a > b : op == O_GE ? a >= b : op == O_EQ ? a == b : a != b;
		} else {
			int64_t a = p->i, b = q->i;
			c = op == O_LT ? a < b : op == O_LE ? a <= b : op == O_GT ? a > b : op == O_GE ? a >= b : op == O_EQ ? a == b : a != b;
		}
		p->i = c; p->r = (double)c; p->vtype = V_INT;
	} }
}

/* kexpr.c:366-399 */
int orc_expr_eval(const orc_expr_t *ke, int64_t *iv, double *rv, int *vtype)
{
	tok_t *st = (tok_t*)malloc((size_t)(ke->n + 1) * sizeof(tok_t));
	int i, top = 0, err = 0;
	for (i = 0; i < ke->n; ++i) {
		const tok_t *e = &ke->e[i];
		if ((e->ttype == T_OP || e->ttype == T_FUNC) && !e->defined) err |= E_UNFUNC;
		else if (e->ttype == T_VAL && e->name && !e->assigned) err |= E_UNVAR;
	}
	for (i = 0; i < ke->n; ++i) {
		const tok_t *e = &ke->e[i];
		if (e->ttype == T_OP || e->ttype == T_FUNC) {
			if (e->n_args == 2 && e->defined) { tok_t *q = &st[--top]; apply(e->op, &st[top-1], q); }
			else if (e->n_args == 1 && e->defined) {
				tok_t *p = &st[top-1];
				if (e->is_abs) { /* kexpr.c:155: abs() takes int */
					if (p->vtype == V_INT) p->i = abs((int)p->i), p->r = (double)p->i;
					else p->r = fabs(p->r), p->i = r2i(p->r);
				} else apply(e->op, p, 0);
			} else top -= e->n_args - 1;
		} else st[top++] = *e;
	}
	*vtype = st[0].vtype; *iv = st[0].i; *rv = st[0].r;
	free(st);
	return err;
}
