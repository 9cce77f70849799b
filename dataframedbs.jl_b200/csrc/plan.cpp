// plan.cpp -- parses the serialized view plan (SelectionQueue + Projection) and compiles every
// BlockBroadcasting tree into (a) a typed, tag-free VM program for the generic kernels and (b), when
// the tree is a conjunction of `column <cmp> constant` terms, the term list of the fused fast path.
//
// Reference objects this stands in for:
//   SelectionQueue / RangeToProcess      /root/reference/src/tables/selection.jl:4-10,68-75
//   _check_element (Bool-only predicate) /root/reference/src/tables/selection.jl:52-55
//   BlockBroadcasting result typing      /root/reference/src/tables/broadcast.jl:6-17 (Base._return_type)
//   Projection                           /root/reference/src/tables/projection.jl:1-9
// Julia semantics implemented by the typing below: exact mixed integer/float comparisons, wrapping
// integer arithmetic with promote_type widths, `/` always floating, three-valued logic on missing.
#include <algorithm>
#include <cmath>
#include <memory>

#include "internal.hpp"

namespace dfdb {
namespace {

struct Rdr {
    const uint8_t *p, *end;
    bool bad = false;
    uint8_t u8() { if (p + 1 > end) { bad = true; return 0; } return *p++; }
    uint32_t u32() { uint32_t v = 0; if (p + 4 > end) { bad = true; return 0; } memcpy(&v, p, 4); p += 4; return v; }
    int64_t i64() { int64_t v = 0; if (p + 8 > end) { bad = true; return 0; } memcpy(&v, p, 8); p += 8; return v; }
    double f64() { double v = 0; if (p + 8 > end) { bad = true; return 0; } memcpy(&v, p, 8); p += 8; return v; }
};

struct Node {
    int op = 0;
    int64_t i = 0;
    double f = 0;
    std::string s;
    std::vector<int64_t> set;
    std::vector<std::unique_ptr<Node>> kids;
};

// static type of a VM stack entry
struct TVal {
    int cls = VC_NONE;   // VC_INT, VC_UINT, VC_FLT, VC_BOOL, VC_STR
    int bits = 64;
    bool nullable = false;
};

int arity(int op)
{
    switch (op) {
    case W_COL: case W_I64: case W_F64: case W_STR: case W_BOOL: return 0;
    case W_NOT: case W_NEG: case W_ISMISSING: case W_IN: return 1;
    case W_EQ: case W_NE: case W_LT: case W_LE: case W_GT: case W_GE:
    case W_AND: case W_OR: case W_XOR:
    case W_ADD: case W_SUB: case W_MUL: case W_DIV: case W_REM:
    case W_COALESCE: case W_STARTSWITH: case W_ENDSWITH: return 2;
    default: return -1;
    }
}

int parse_tree(Rdr &r, std::unique_ptr<Node> *out)
{
    uint32_t nops = r.u32();
    if (r.bad || nops == 0 || nops > 4096) return fail(DFDB_ERR_ARGUMENT, "bad expression");
    std::vector<std::unique_ptr<Node>> st;
    for (uint32_t k = 0; k < nops; k++) {
        auto n = std::make_unique<Node>();
        n->op = r.u8();
        switch (n->op) {
        case W_COL: case W_I64: n->i = r.i64(); break;
        case W_F64: n->f = r.f64(); break;
        case W_BOOL: n->i = r.u8() != 0; break;
        case W_STR: {
            uint32_t len = r.u32();
            if (r.bad || r.p + len > r.end) return fail(DFDB_ERR_ARGUMENT, "truncated expression");
            n->s.assign(reinterpret_cast<const char *>(r.p), len);
            r.p += len;
            break;
        }
        case W_IN: {
            uint32_t cnt = r.u32();
            if (r.bad || r.p + 8ull * cnt > r.end) return fail(DFDB_ERR_ARGUMENT, "truncated expression");
            n->set.resize(cnt);
            if (cnt) memcpy(n->set.data(), r.p, 8ull * cnt);
            r.p += 8ull * cnt;
            break;
        }
        default: break;
        }
        if (r.bad) return fail(DFDB_ERR_ARGUMENT, "truncated expression");
        int ar = arity(n->op);
        if (ar < 0) return fail(DFDB_ERR_UNSUPPORTED, "unknown opcode 0x%02x", n->op);
        if ((int)st.size() < ar) return fail(DFDB_ERR_ARGUMENT, "malformed expression (stack underflow)");
        n->kids.resize((size_t)ar);
        for (int a = ar - 1; a >= 0; a--) { n->kids[(size_t)a] = std::move(st.back()); st.pop_back(); }
        st.push_back(std::move(n));
    }
    if (st.size() != 1) return fail(DFDB_ERR_ARGUMENT, "malformed expression");
    *out = std::move(st[0]);
    return DFDB_OK;
}

struct Compiler {
    dfdb_table *tbl;
    dfdb_scan *scan;
    Expr *e;
    VmProgram *prog;
    int nconst = 0, npool = 0;
    int depth = 0, maxdepth = 0;

    int slot_of(int64_t id)
    {
        for (size_t i = 0; i < scan->slots.size(); i++) if (scan->slots[i] == id) return (int)i;
        scan->slots.push_back(id);
        return (int)scan->slots.size() - 1;
    }
    int emit(uint8_t op, uint8_t a = 0, uint8_t b = 0, uint8_t c = 0, int32_t imm = 0)
    {
        if (prog->ninstr >= VM_MAX_INSTR) return fail(DFDB_ERR_UNSUPPORTED, "expression too long");
        prog->instr[prog->ninstr++] = VmInstr{op, a, b, c, imm};
        return DFDB_OK;
    }
    int add_const(int64_t bits, int *idx)
    {
        if (nconst >= VM_MAX_CONST) return fail(DFDB_ERR_UNSUPPORTED, "too many constants in expression");
        prog->consts[nconst] = bits;
        *idx = nconst++;
        return DFDB_OK;
    }
    void push() { if (++depth > maxdepth) maxdepth = depth; }

    static int cls_code(const TVal &v) { return v.cls == VC_FLT ? 3 : v.cls == VC_UINT ? 2 : v.cls == VC_BOOL ? 4 : 1; }

    int compile(const Node *n, TVal *out)
    {
        int rc;
        switch (n->op) {
        case W_COL: {
            Column *c = tbl->find(n->i);
            if (!c) return fail(DFDB_ERR_KEY, "unknown column id %lld", (long long)n->i);
            int cls = value_class(c->type.kind);
            if (cls == VC_NONE) return fail(DFDB_ERR_UNSUPPORTED, "column %s of type %s is not supported in expressions", c->name.c_str(), c->typestr.c_str());
            int slot = slot_of(n->i);
            if (slot >= MAX_SLOTS) return fail(DFDB_ERR_UNSUPPORTED, "too many columns in one scan (max %d)", MAX_SLOTS);
            if (std::find(e->col_ids.begin(), e->col_ids.end(), n->i) == e->col_ids.end()) e->col_ids.push_back(n->i);
            out->cls = cls;
            out->bits = cls == VC_STR ? 0 : c->type.elsize * 8;
            out->nullable = c->type.nullable;
            push();
            return emit(V_LOAD, (uint8_t)slot);
        }
        case W_I64: {
            int idx;
            if ((rc = add_const(n->i, &idx))) return rc;
            *out = TVal{VC_INT, 64, false};
            push();
            return emit(V_CONST, 0, 0, 0, idx);
        }
        case W_F64: {
            int idx;
            int64_t bits;
            memcpy(&bits, &n->f, 8);
            if ((rc = add_const(bits, &idx))) return rc;
            *out = TVal{VC_FLT, 64, false};
            push();
            return emit(V_CONST, 0, 0, 0, idx);
        }
        case W_BOOL: {
            int idx;
            if ((rc = add_const(n->i ? 1 : 0, &idx))) return rc;
            *out = TVal{VC_BOOL, 8, false};
            push();
            return emit(V_CONST, 0, 0, 0, idx);
        }
        case W_STR: {
            if (npool + (int)n->s.size() > VM_STRPOOL || n->s.size() > 0xFFFF) return fail(DFDB_ERR_UNSUPPORTED, "string constants too long");
            int off = npool;
            memcpy(prog->strpool + npool, n->s.data(), n->s.size());
            npool += (int)n->s.size();
            *out = TVal{VC_STR, 0, false};
            push();
            return emit(V_CONSTSTR, (uint8_t)(n->s.size() & 0xFF), (uint8_t)(n->s.size() >> 8), 0, off);
        }
        default: break;
        }
        TVal a, b;
        if ((rc = compile(n->kids[0].get(), &a))) return rc;
        if (n->kids.size() > 1 && (rc = compile(n->kids[1].get(), &b))) return rc;
        bool nullable = a.nullable || (n->kids.size() > 1 && b.nullable);
        switch (n->op) {
        case W_EQ: case W_NE: case W_LT: case W_LE: case W_GT: case W_GE: {
            int code = n->op - W_EQ;
            depth--;
            *out = TVal{VC_BOOL, 8, nullable};
            if (a.cls == VC_STR || b.cls == VC_STR) {
                if (a.cls == b.cls) return emit(V_STRCMP, (uint8_t)code);
                // ==(::String, ::Number) is false in Julia; ordering is a MethodError
                if (code > 1) return fail(DFDB_ERR_UNSUPPORTED, "ordering comparison between String and non-String");
                return emit(V_STRNUM, (uint8_t)code);
            }
            int la = a.cls == VC_FLT ? 2 : a.cls == VC_UINT ? 1 : 0, lb = b.cls == VC_FLT ? 2 : b.cls == VC_UINT ? 1 : 0;
            static const int cp[3][3] = {{CP_II, CP_IU, CP_IF}, {CP_UI, CP_UU, CP_UF}, {CP_FI, CP_FU, CP_FF}};
            return emit(V_CMP, (uint8_t)code, (uint8_t)cp[la][lb]);
        }
        case W_AND: case W_OR: case W_XOR:
            if (a.cls != VC_BOOL || b.cls != VC_BOOL) return fail(DFDB_ERR_UNSUPPORTED, "bitwise logic is supported on Bool operands only");
            depth--;
            *out = TVal{VC_BOOL, 8, nullable};
            return emit(n->op == W_AND ? V_AND : n->op == W_OR ? V_OR : V_XOR);
        case W_NOT:
            if (a.cls != VC_BOOL) return fail(DFDB_ERR_UNSUPPORTED, "! is supported on Bool operands only");
            *out = TVal{VC_BOOL, 8, nullable};
            return emit(V_NOT);
        case W_ADD: case W_SUB: case W_MUL: case W_DIV: case W_REM: case W_NEG: {
            bool unary = n->op == W_NEG;
            if (a.cls == VC_STR || (!unary && b.cls == VC_STR)) return fail(DFDB_ERR_UNSUPPORTED, "arithmetic on String");
            int code = n->op - W_ADD;
            if (!unary) depth--;
            bool af = a.cls == VC_FLT, bf = !unary && b.cls == VC_FLT;
            uint8_t classes = (uint8_t)(cls_code(a) | ((unary ? 0 : cls_code(b)) << 4));
            if (n->op == W_DIV || af || bf) {
                int bits = 64;
                if (!(n->op == W_DIV && !af && !bf)) {
                    int fa = af ? a.bits : 0, fb = bf ? b.bits : 0;
                    bits = (fa == 64 || fb == 64) ? 64 : 32;
                }
                *out = TVal{VC_FLT, bits, nullable};
                return emit(V_ARITH, (uint8_t)code, classes, (uint8_t)(0x40 | (bits / 8)));
            }
            int abits = (a.cls == VC_INT || a.cls == VC_UINT) ? a.bits : 0;
            int bbits = (!unary && (b.cls == VC_INT || b.cls == VC_UINT)) ? b.bits : 0;
            bool auns = a.cls == VC_UINT, buns = !unary && b.cls == VC_UINT;
            int bits;
            bool uns;
            if (unary) { bits = abits ? abits : 64; uns = auns; }
            else if (!abits && !bbits) { bits = 64; uns = false; }
            else if (!abits) { bits = bbits; uns = buns; }
            else if (!bbits) { bits = abits; uns = auns; }
            else if (abits == bbits) { bits = abits; uns = auns || buns; }
            else if (abits > bbits) { bits = abits; uns = auns; }
            else { bits = bbits; uns = buns; }
            *out = TVal{uns ? VC_UINT : VC_INT, bits, nullable};
            return emit(V_ARITH, (uint8_t)code, classes, (uint8_t)((uns ? 0x80 : 0) | (bits / 8)));
        }
        case W_ISMISSING:
            *out = TVal{VC_BOOL, 8, false};
            return emit(V_ISMISSING);
        case W_COALESCE: {
            bool same = (a.cls == b.cls) || ((a.cls == VC_INT || a.cls == VC_UINT) && (b.cls == VC_INT || b.cls == VC_UINT));
            if (!same || n->kids[1]->kids.size() != 0 || n->kids[1]->op == W_COL)
                return fail(DFDB_ERR_UNSUPPORTED, "coalesce default must be a constant of the same kind");
            depth--;
            *out = a;
            out->nullable = false;
            return emit(V_COALESCE);
        }
        case W_STARTSWITH: case W_ENDSWITH:
            if (a.cls != VC_STR || b.cls != VC_STR) return fail(DFDB_ERR_UNSUPPORTED, "startswith/endswith need String operands");
            depth--;
            *out = TVal{VC_BOOL, 8, nullable};
            return emit(n->op == W_STARTSWITH ? V_STARTSWITH : V_ENDSWITH);
        case W_IN: {
            if (a.cls != VC_INT && a.cls != VC_UINT && a.cls != VC_BOOL) return fail(DFDB_ERR_UNSUPPORTED, "in() is supported for integer values");
            if (n->set.size() > 0xFFFF) return fail(DFDB_ERR_UNSUPPORTED, "in() set too large");
            int first = nconst;
            for (int64_t v : n->set) {
                int idx;
                if ((rc = add_const(v, &idx))) return rc;
            }
            *out = TVal{VC_BOOL, 8, a.nullable};
            return emit(V_IN, (uint8_t)(n->set.size() & 0xFF), (uint8_t)(n->set.size() >> 8), (uint8_t)(a.cls == VC_UINT), first);
        }
        default:
            return fail(DFDB_ERR_UNSUPPORTED, "unknown opcode 0x%02x", n->op);
        }
    }
};

// ---- fast path matching ------------------------------------------------------------------------

bool is_numeric_const(const Node *n) { return n->op == W_I64 || n->op == W_F64 || n->op == W_BOOL; }

int flip_code(int code)
{
    switch (code) {
    case 2: return 4;   // LT -> GT
    case 3: return 5;   // LE -> GE
    case 4: return 2;
    case 5: return 3;
    default: return code;
    }
}

// result of `x <code> c` when the comparison is decided by c alone (x below / above c)
int const_cmp(int code, int sign)   // sign = sgn(x - c): -1 means every x is < c
{
    switch (code) {
    case 0: return 0;
    case 1: return 1;
    case 2: case 3: return sign < 0;
    default: return sign > 0;
    }
}

bool make_term(dfdb_table *tbl, dfdb_scan *scan, const Node *col, const Node *cst, int code, Term *t)
{
    Column *c = tbl->find(col->i);
    if (!c) return false;
    int cls = value_class(c->type.kind);
    if (cls != VC_INT && cls != VC_UINT && cls != VC_FLT && cls != VC_BOOL) return false;
    int slot = -1;
    for (size_t i = 0; i < scan->slots.size(); i++) if (scan->slots[i] == col->i) slot = (int)i;
    if (slot < 0) return false;
    t->slot = slot;
    t->code = code;
    t->constant_result = -1;
    t->ci = 0;
    t->cf = 0;
    if (cls == VC_BOOL) cls = VC_UINT;
    t->cls = cls;
    bool cf_const = cst->op == W_F64;
    int64_t ci = cst->i;
    double cf = cst->f;
    if (cls == VC_FLT) {
        if (cf_const) { t->cf = cf; return true; }
        double d = (double)ci;
        bool exact = d >= -9223372036854775808.0 && d < 9223372036854775808.0 && (int64_t)d == ci;
        if (exact) { t->cf = d; return true; }
        // ci lies strictly between two adjacent doubles
        bool d_above = d >= 9223372036854775808.0 || (d > -9223372036854775808.0 && (__int128)(int64_t)d > (__int128)ci);
        double lo = d_above ? std::nextafter(d, -INFINITY) : d, hi = d_above ? d : std::nextafter(d, INFINITY);
        switch (code) {
        case 0: t->constant_result = 0; break;
        case 1: t->constant_result = 1; break;   // NaN != c is true as well
        case 2: case 3: t->code = 3; t->cf = lo; break;
        default: t->code = 5; t->cf = hi; break;
        }
        return true;
    }
    if (!cf_const) {
        if (cls == VC_UINT) {
            if (ci < 0) { t->constant_result = const_cmp(code, +1); return true; }
            t->ci = ci;
            return true;
        }
        t->ci = ci;
        return true;
    }
    // integer column against a Float64 constant: fold to an exact integer comparison
    if (cf != cf) { t->constant_result = code == 1; return true; }
    double lo_lim = cls == VC_UINT ? 0.0 : -9223372036854775808.0;
    double hi_lim = cls == VC_UINT ? 18446744073709551616.0 : 9223372036854775808.0;
    if (cf >= hi_lim) { t->constant_result = const_cmp(code, -1); return true; }
    if (cf < lo_lim) { t->constant_result = const_cmp(code, +1); return true; }
    double fl = std::floor(cf);
    int64_t k = cls == VC_UINT ? (int64_t)(uint64_t)fl : (int64_t)fl;
    t->ci = k;
    if (fl == cf) return true;
    switch (code) {
    case 0: t->constant_result = 0; break;
    case 1: t->constant_result = 1; break;
    case 2: case 3: t->code = 3; break;   // x < c  <=>  x <= floor(c)
    default: t->code = 4; break;          // x > c, x >= c  <=>  x > floor(c)
    }
    return true;
}

bool match_simple(dfdb_table *tbl, dfdb_scan *scan, const Node *n, bool in_coalesce, Expr *e)
{
    if (n->op == W_AND) return match_simple(tbl, scan, n->kids[0].get(), in_coalesce, e) && match_simple(tbl, scan, n->kids[1].get(), in_coalesce, e);
    if (n->op == W_COALESCE) {
        const Node *d = n->kids[1].get();
        if (d->op != W_BOOL || d->i != 0) return false;
        return match_simple(tbl, scan, n->kids[0].get(), true, e);
    }
    if (n->op >= W_EQ && n->op <= W_GE) {
        const Node *l = n->kids[0].get(), *r = n->kids[1].get();
        int code = n->op - W_EQ;
        const Node *col, *cst;
        if (l->op == W_COL && is_numeric_const(r)) { col = l; cst = r; }
        else if (r->op == W_COL && is_numeric_const(l)) { col = r; cst = l; code = flip_code(code); }
        else return false;
        Column *c = tbl->find(col->i);
        if (!c || (c->type.nullable && !in_coalesce)) return false;
        if (e->nterms >= MAX_TERMS) return false;
        Node tmp;
        if (cst->op == W_BOOL) { tmp.op = W_I64; tmp.i = cst->i; cst = &tmp; }
        if (!make_term(tbl, scan, col, cst, code, &e->terms[e->nterms])) return false;
        e->nterms++;
        return true;
    }
    return false;
}

int compile_expr(dfdb_table *tbl, dfdb_scan *scan, Rdr &r, Expr *e, TVal *result)
{
    const uint8_t *start = r.p;
    std::unique_ptr<Node> root;
    int rc = parse_tree(r, &root);
    if (rc) return rc;
    e->wire.assign(start, r.p);
    memset(&e->prog, 0, sizeof e->prog);
    Compiler c{tbl, scan, e, &e->prog};
    rc = c.compile(root.get(), result);
    if (rc) return rc;
    if (c.maxdepth > VM_MAX_STACK) return fail(DFDB_ERR_UNSUPPORTED, "expression too deep");
    e->prog.result_class = result->cls;
    e->prog.result_nullable = result->nullable;
    e->prog.result_bits = result->bits;
    e->prog.result_uns = result->cls == VC_UINT;
    e->nterms = 0;
    e->simple = result->cls == VC_BOOL && !result->nullable && match_simple(tbl, scan, root.get(), false, e);
    if (!e->simple) e->nterms = 0;
    return DFDB_OK;
}

}  // namespace

int plan_parse(dfdb_table *t, const uint8_t *bytes, int64_t len, dfdb_scan *s)
{
    Rdr r{bytes, bytes + len};
    if (len < 12 || r.u32() != 0x31504644u) return fail(DFDB_ERR_ARGUMENT, "bad plan magic");
    uint32_t nstages = r.u32();
    if (r.bad || nstages > 1024) return fail(DFDB_ERR_ARGUMENT, "bad plan");
    s->stages.resize(nstages);
    for (auto &st : s->stages) {
        st.kind = r.u8();
        if (st.kind == ST_RANGE) {
            st.start = r.i64(); st.step = r.i64(); st.stop = r.i64();
            if (r.bad) return fail(DFDB_ERR_ARGUMENT, "truncated plan");
            if (st.step == 0) return fail(DFDB_ERR_ARGUMENT, "step cannot be zero");
            bool empty = st.step > 0 ? st.start > st.stop : st.start < st.stop;
            // RangeToProcess takes minimum/maximum of the range, which throw on an empty range (selection.jl:73)
            if (empty) return fail(DFDB_ERR_ARGUMENT, "range must be non-empty");
            st.first = std::min(st.start, st.stop);
            st.last = std::max(st.start, st.stop);
        } else if (st.kind == ST_INDEXVEC) {
            uint32_t n = r.u32();
            if (r.bad || r.p + 8ull * n > r.end) return fail(DFDB_ERR_ARGUMENT, "truncated plan");
            if (n == 0) return fail(DFDB_ERR_ARGUMENT, "reducing over an empty collection is not allowed");
            st.idx.resize(n);
            memcpy(st.idx.data(), r.p, 8ull * n);
            r.p += 8ull * n;
            std::sort(st.idx.begin(), st.idx.end());
            st.idx.erase(std::unique(st.idx.begin(), st.idx.end()), st.idx.end());
            st.first = st.idx.front();
            st.last = st.idx.back();
        } else if (st.kind == ST_PRED) {
            TVal res;
            int rc = compile_expr(t, s, r, &st.e, &res);
            if (rc) return rc;
            // _check_element selection.jl:52-55
            if (res.cls != VC_BOOL || res.nullable) return fail(DFDB_ERR_ARGUMENT, "Function for selection must have Bool result type");
        } else {
            return fail(DFDB_ERR_ARGUMENT, "bad stage kind %d", st.kind);
        }
    }
    uint32_t nproj = r.u32();
    if (r.bad || nproj > 4096) return fail(DFDB_ERR_ARGUMENT, "bad plan");
    s->projs.resize(nproj);
    for (auto &p : s->projs) {
        p.kind = r.u8();
        if (p.kind == PJ_COL) {
            p.col = r.i64();
            if (r.bad) return fail(DFDB_ERR_ARGUMENT, "truncated plan");
            Column *c = t->find(p.col);
            if (!c) return fail(DFDB_ERR_KEY, "unknown column id %lld", (long long)p.col);
            p.type = c->type;
        } else if (p.kind == PJ_EXPR) {
            TVal res;
            int rc = compile_expr(t, s, r, &p.e, &res);
            if (rc) return rc;
            p.type.nullable = res.nullable;
            if (res.cls == VC_INT || res.cls == VC_UINT) {
                static const int sk[] = {DFDB_I8, DFDB_I16, DFDB_I32, DFDB_I64}, uk[] = {DFDB_U8, DFDB_U16, DFDB_U32, DFDB_U64};
                int w = res.bits == 8 ? 0 : res.bits == 16 ? 1 : res.bits == 32 ? 2 : 3;
                p.type.kind = res.cls == VC_UINT ? uk[w] : sk[w];
                p.type.elsize = res.bits / 8;
            } else if (res.cls == VC_FLT) {
                p.type.kind = res.bits == 32 ? DFDB_F32 : DFDB_F64;
                p.type.elsize = res.bits / 8;
            } else if (res.cls == VC_BOOL) {
                p.type.kind = DFDB_BOOL;
                p.type.elsize = 1;
            } else {
                return fail(DFDB_ERR_UNSUPPORTED, "computed String columns are not supported");
            }
        } else {
            return fail(DFDB_ERR_ARGUMENT, "bad projection kind %d", p.kind);
        }
    }
    if (s->slots.size() > (size_t)MAX_SLOTS) return fail(DFDB_ERR_UNSUPPORTED, "too many columns in one scan (max %d)", MAX_SLOTS);
    return DFDB_OK;
}

}  // namespace dfdb
