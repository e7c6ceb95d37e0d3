// flt.cpp -- host-side compiler of `bgt view -f` site filters into flt_prog_t byte-code.
//
// Grammar, precedence and error codes are those of kexpr (kexpr.c:62-76 precedence table, :163-244 tokens,
// :257-352 shunting-yard, kexpr.h:10-16 error bits); literals go through the same libc strtod/strtol(base 0)
// pair the reference uses (kexpr.c:181-196), so "010" is the integer 8 with real part 10.0 here as well.
// Variables are resolved at compile time against the names bgtm_assign_expr binds (bgt.c:700-710):
// AN, AC, AN<g>, AC<g> for g = 1..n_groups; anything else stays unbound and fails every site (kexpr.c:373).
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "flt.h"

namespace {

enum { PE_UNQU = 0x01, PE_UNLP = 0x02, PE_UNRP = 0x04, PE_UNOP = 0x08, PE_FUNC = 0x10, PE_ARG = 0x20, PE_NUM = 0x40,
       PE_TOO_COMPLEX = 0x100 };

struct Tok {
	enum Kind { VAL, OP, FUNC, LPAREN } kind = VAL;
	int op = 0, n_args = 0, vtype = 0;
	std::string name, str;
	int64_t i = 0;
	double r = 0;
};

const int kPrec[26] = {0, 1,1,1,1, 2, 3,3,3,3, 4,4, 5,5, 6,6,6,6, 7,7, 8, 9, 10, 11, 12, 0};
inline bool right_assoc(int op) { return op >= FO_POS && op <= FO_POW; }

bool lex(const char *&p, bool last_is_val, Tok &t, int &err)
{
	const char *q = p;
	if (isalpha((unsigned char)*p) || *p == '_') {
		while (*p && (*p == '_' || isalnum((unsigned char)*p))) ++p;
		t.name.assign(q, p - q);
		if (*p == '(') t.kind = Tok::FUNC, t.n_args = 1;
		else t.kind = Tok::VAL, t.vtype = FV_REAL;
		return true;
	}
	if (isdigit((unsigned char)*p) || *p == '.') {
		char *pd, *pl;
		const double y = strtod(q, &pd);
		const long x = strtol(q, &pl, 0);
		t.kind = Tok::VAL;
		if (pd == q && pl == q) { err |= PE_NUM; return false; }
		if (pd > pl) t.vtype = FV_REAL, t.r = y, t.i = flt_r2i(y), p = pd;
		else t.vtype = FV_INT, t.r = y, t.i = x, p = pl;
		return true;
	}
	if (*p == '"' || *p == '\'') {
		const char c = *p;
		for (++p; *p && *p != c; ++p) if (*p == '\\') ++p;
		if (*p != c) { err |= PE_UNQU; return false; }
		t.kind = Tok::VAL; t.vtype = FV_STR; t.str.assign(q + 1, p - q - 1);
		++p;
		return true;
	}
	t.kind = Tok::OP; t.n_args = 2;
	auto two = [&](char a, char b) { return p[0] == a && p[1] == b; };
	if (two('*', '*')) t.op = FO_POW, p += 2;
	else if (*p == '*') t.op = FO_MUL, ++p;
	else if (two('/', '/')) t.op = FO_IDIV, p += 2;
	else if (*p == '/') t.op = FO_DIV, ++p;
	else if (*p == '%') t.op = FO_MOD, ++p;
	else if (*p == '+') { t.op = last_is_val ? FO_ADD : FO_POS; t.n_args = last_is_val ? 2 : 1; ++p; }
	else if (*p == '-') { t.op = last_is_val ? FO_SUB : FO_NEG; t.n_args = last_is_val ? 2 : 1; ++p; }
	else if (two('=', '=')) t.op = FO_EQ, p += 2;
	else if (two('!', '=') || two('<', '>')) t.op = FO_NE, p += 2;
	else if (two('>', '=')) t.op = FO_GE, p += 2;
	else if (two('<', '=')) t.op = FO_LE, p += 2;
	else if (two('>', '>')) t.op = FO_RSH, p += 2;
	else if (two('<', '<')) t.op = FO_LSH, p += 2;
	else if (*p == '>') t.op = FO_GT, ++p;
	else if (*p == '<') t.op = FO_LT, ++p;
	else if (two('|', '|')) t.op = FO_LOR, p += 2;
	else if (two('&', '&')) t.op = FO_LAND, p += 2;
	else if (*p == '|') t.op = FO_BOR, ++p;
	else if (*p == '&') t.op = FO_BAND, ++p;
	else if (*p == '^') t.op = FO_BXOR, ++p;
	else if (*p == '~') t.op = FO_BNOT, t.n_args = 1, ++p;
	else if (*p == '!') t.op = FO_LNOT, t.n_args = 1, ++p;
	else { err |= PE_UNOP; return false; }
	return true;
}

// slot of a bound variable in the per-site count vector [AN, AC, AC<M>, AN1, AC1, AC1<M>, ...], or -1
int var_slot(const std::string &name, int n_groups)
{
	if (name.size() < 2 || name[0] != 'A' || (name[1] != 'N' && name[1] != 'C')) return -1;
	const bool is_n = name[1] == 'N';
	if (name.size() == 2) return is_n ? 0 : 1;
	// bgt.c:692-698: one digit for groups 1..9, two for 10..32; no leading zero forms are ever bound
	int g = 0;
	for (size_t k = 2; k < name.size(); ++k) {
		if (!isdigit((unsigned char)name[k])) return -1;
		g = g * 10 + (name[k] - '0');
	}
	if (name.size() > 4 || name[2] == '0') return -1;
	if ((name.size() == 3 && (g < 1 || g > 9)) || (name.size() == 4 && g < 10)) return -1;
	if (g < 1 || g > n_groups) return -1;
	return 3 + 3 * (g - 1) + (is_n ? 0 : 1);
}

} // namespace

int flt_compile(const char *expr, int n_groups, flt_prog_t *P)
{
	memset(P, 0, sizeof(*P));
	if (expr == 0) return 0;
	std::string s;
	for (const char *c = expr; *c; ++c) if (!isspace((unsigned char)*c)) s.push_back(*c);
	std::vector<Tok> out, ops;
	int err = 0;
	bool last_is_val = false;
	const char *p = s.c_str();
	auto pop_until_paren = [&]() { while (!ops.empty() && ops.back().kind != Tok::LPAREN) { out.push_back(ops.back()); ops.pop_back(); } };
	while (*p) {
		if (*p == '(') { Tok t; t.kind = Tok::LPAREN; ops.push_back(t); ++p; }
		else if (*p == ')') {
			pop_until_paren();
			if (ops.empty()) { err |= PE_UNRP; break; }
			ops.pop_back();
			if (!ops.empty() && ops.back().kind == Tok::FUNC) { out.push_back(ops.back()); ops.pop_back(); }
			++p;
		} else if (*p == ',') {
			pop_until_paren();
			if (ops.size() < 2 || ops[ops.size() - 2].kind != Tok::FUNC) { err |= PE_FUNC; break; }
			++ops[ops.size() - 2].n_args; ++p;
		} else {
			Tok t;
			if (!lex(p, last_is_val, t, err)) break;
			if (t.kind == Tok::VAL) out.push_back(t), last_is_val = true;
			else if (t.kind == Tok::FUNC) ops.push_back(t), last_is_val = false;
			else {
				while (!ops.empty() && ops.back().kind == Tok::OP) {
					const int top = kPrec[ops.back().op], me = kPrec[t.op];
					if (right_assoc(t.op) ? me <= top : me < top) break;
					out.push_back(ops.back()); ops.pop_back();
				}
				ops.push_back(t); last_is_val = false;
			}
		}
	}
	if (!err) { pop_until_paren(); if (!ops.empty()) err |= PE_UNLP; }
	if (!err) {
		int n = 0;
		for (const Tok &t : out) n += t.kind == Tok::VAL ? 1 : -(t.n_args - 1);
		if (n != 1) err |= PE_ARG;
	}
	if (err) return err;

	// lower to byte-code
	std::vector<std::string> strs;
	int depth = 0, max_depth = 0;
	if (out.size() > FLT_MAX_CODE) return PE_TOO_COMPLEX;
	for (const Tok &t : out) {
		flt_ins_t &e = P->code[P->n++];
		if (t.kind == Tok::VAL) {
			if (!t.name.empty()) {
				const int slot = var_slot(t.name, n_groups);
				if (slot < 0) { P->always_fail = 1; e.kind = FK_CONST; e.vtype = FV_REAL; }
				else { e.kind = FK_VAR; e.arg = slot; }
			} else {
				e.kind = FK_CONST; e.vtype = (uint8_t)t.vtype; e.i = t.i; e.r = t.r;
				if (t.vtype == FV_STR) {
					size_t k = 0;
					while (k < strs.size() && strs[k] != t.str) ++k;
					if (k == strs.size()) strs.push_back(t.str);
					if (strs.size() > FLT_MAX_STR) return PE_TOO_COMPLEX;
					e.sid = (uint8_t)k;
				}
			}
			++depth;
		} else if (t.kind == Tok::OP) {
			e.kind = t.n_args == 2 ? FK_OP2 : FK_OP1; e.op = (uint8_t)t.op;
			if (t.op == FO_POW) P->needs_host = 1;
			depth -= t.n_args - 1;
		} else { // function call: only abs(x) has an implementation for site filters (kexpr.c:286; bgt.c:444-455)
			if (t.n_args == 1 && t.name == "abs") e.kind = FK_OP1, e.op = FO_ABS;
			else { P->always_fail = 1; e.kind = FK_DROP; e.arg = t.n_args - 1; depth -= t.n_args - 1; }
		}
		if (depth < 1) P->always_fail = 1; // malformed operand order; the reference would read below its stack
		if (depth > max_depth) max_depth = depth;
	}
	if (max_depth > FLT_MAX_STACK) return PE_TOO_COMPLEX;
	P->n_str = (int)strs.size();
	for (size_t a = 0; a < strs.size(); ++a)
		for (size_t b = 0; b < strs.size(); ++b) {
			const int d = strcmp(strs[a].c_str(), strs[b].c_str());
			P->scmp[a][b] = (int8_t)((d > 0) - (d < 0));
		}
	return 0;
}
