/*
 * oracle/dfdb_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Single-threaded CPU restatement of the column-scan hot path of DataFrameDBs.jl, used as the
 * parity checker for the CUDA library (tests/, __graft_entry__.smoke(), bench.py cpu_baseline /
 * --impl reference).  It is never imported, linked or executed by the product path.
 *
 * Parity status: the reference is Julia and cannot run in this image (no julia binary), so this
 * file is pinned against (a) every known-answer case of the reference's own tests for this path
 * (tests/test_oracle_golden.py restates /root/reference/test/{selection,missings,flat_strings,
 * block_streams,view,columnbroadcast,range_indexing,column}.jl with independent numpy
 * expectations), (b) the documented LZ4 ratios, and (c) the system liblz4 for the codec.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdarg.h>

#define ORC_API __attribute__((visibility("default")))

int orc_lz4_decompress_safe(const uint8_t *src, uint8_t *dst, int srcSize, int dstCap);

/* ------------------------------------------------------------------------------------------ */
/* errors                                                                                      */

enum {
    ORC_OK = 0,
    ORC_ERR_IO = 1,          /* error(...) in filesystem.jl / creators.jl                     */
    ORC_ERR_FORMAT = 2,      /* header mismatch filesystem.jl:47-54                           */
    ORC_ERR_CORRUPT = 3,     /* @assert size == sizes.origin  BlockStreams.jl:112             */
    ORC_ERR_ARGUMENT = 4,    /* ArgumentError (selection.jl:54, empty range, ...)             */
    ORC_ERR_UNSUPPORTED = 5, /* expression outside the supported operator set                 */
    ORC_ERR_KEY = 6,         /* unknown column                                                */
    ORC_ERR_DIVIDE = 7       /* DivideError from integer rem/div by zero                      */
};

static __thread char g_err[512];
ORC_API const char *orc_last_error(void) { return g_err; }
static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

/* ------------------------------------------------------------------------------------------ */
/* column types: src/columntypes/base.jl:41-74,97-126,163-168 ; complex.jl:1-20                */

enum {
    K_I8 = 1, K_I16, K_I32, K_I64, K_I128, K_U8, K_U16, K_U32, K_U64, K_U128,
    K_F16, K_F32, K_F64, K_BOOL, K_CHAR, K_STRING, K_DATE, K_DATETIME, K_TIME, K_TUPLE
};

typedef struct {
    int kind;
    int nullable;
    int elsize;   /* bytes per element in a block body (0 for strings) */
    int align;
} coltype;

static const struct { const char *name; int kind, size; } k_prims[] = {
    {"Int8", K_I8, 1}, {"Int16", K_I16, 2}, {"Int32", K_I32, 4}, {"Int64", K_I64, 8}, {"Int128", K_I128, 16},
    {"UInt8", K_U8, 1}, {"UInt16", K_U16, 2}, {"UInt32", K_U32, 4}, {"UInt64", K_U64, 8}, {"UInt128", K_U128, 16},
    {"Float16", K_F16, 2}, {"Float32", K_F32, 4}, {"Float64", K_F64, 8}, {"Bool", K_BOOL, 1}, {"Char", K_CHAR, 4},
    {"String", K_STRING, 0}, {"Date", K_DATE, 8}, {"DateTime", K_DATETIME, 8}, {"Time", K_TIME, 8},
};

static void trim(const char **s, size_t *n)
{
    while (*n && ((*s)[0] == ' ' || (*s)[0] == '\t')) { (*s)++; (*n)--; }
    while (*n && ((*s)[*n - 1] == ' ' || (*s)[*n - 1] == '\t')) (*n)--;
}

/* parse_typestring + deserialize in one go; Tuple layout follows Julia's C-compatible struct
 * layout (natural alignment, size rounded to max alignment). */
static int parse_type(const char *s, size_t n, coltype *out)
{
    trim(&s, &n);
    if (n == 0 || s[0] == '(') return fail(ORC_ERR_FORMAT, "typename parse error");
    const char *brace = memchr(s, '(', n);
    if (!brace) {
        for (size_t i = 0; i < sizeof k_prims / sizeof k_prims[0]; i++) {
            if (strlen(k_prims[i].name) == n && memcmp(k_prims[i].name, s, n) == 0) {
                out->kind = k_prims[i].kind;
                out->nullable = 0;
                out->elsize = k_prims[i].size;
                out->align = k_prims[i].size > 8 ? 16 : (k_prims[i].size ? k_prims[i].size : 1);
                return ORC_OK;
            }
        }
        return fail(ORC_ERR_FORMAT, "Undefined column type: %.*s", (int)n, s);
    }
    if (s[n - 1] != ')') return fail(ORC_ERR_FORMAT, "typename parse error");
    size_t hn = (size_t)(brace - s);
    const char *inner = brace + 1;
    size_t in = n - hn - 2;
    if (hn == 7 && memcmp(s, "Missing", 7) == 0) {
        int rc = parse_type(inner, in, out);
        if (rc) return rc;
        if (out->nullable) return fail(ORC_ERR_FORMAT, "nested Missing");
        out->nullable = 1;
        return ORC_OK;
    }
    if (hn == 5 && memcmp(s, "Tuple", 5) == 0) {
        int depth = 0, size = 0, maxal = 1, count = 0;
        size_t start = 0;
        for (size_t i = 0; i <= in; i++) {
            char c = i < in ? inner[i] : ',';
            if (c == '(') depth++;
            else if (c == ')') depth--;
            else if (c == ',' && depth == 0) {
                coltype e;
                if (i > start) {
                    int rc = parse_type(inner + start, i - start, &e);
                    if (rc) return rc;
                    if (e.kind == K_STRING || e.nullable) return fail(ORC_ERR_FORMAT, "Unsupported tuple element");
                    size = (size + e.align - 1) / e.align * e.align + e.elsize;
                    if (e.align > maxal) maxal = e.align;
                    count++;
                }
                start = i + 1;
            }
        }
        if (!count) return fail(ORC_ERR_FORMAT, "Undefined column type: Tuple");
        out->kind = K_TUPLE;
        out->nullable = 0;
        out->align = maxal;
        out->elsize = (size + maxal - 1) / maxal * maxal;
        return ORC_OK;
    }
    return fail(ORC_ERR_FORMAT, "Undefined column type: %.*s", (int)n, s);
}

ORC_API int orc_parse_type(const char *s, int *kind, int *nullable, int *elsize)
{
    coltype t;
    int rc = parse_type(s, strlen(s), &t);
    if (rc) return rc;
    *kind = t.kind; *nullable = t.nullable; *elsize = t.elsize;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* table meta: src/io/table_io.jl:21-33, common_io.jl:5-8, filesystem.jl:8-12,47-54            */

typedef struct {
    int64_t id;
    char name[256];
    char typestr[256];
    coltype type;
    int64_t data_start;   /* byte offset of the first block in <id>.bin */
} colmeta;

typedef struct orc_table {
    char path[1024];
    int64_t format_version, block_size, ncols;
    colmeta *cols;
} orc_table;

static int rd_exact(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n ? 0 : -1; }
static int rd_string(FILE *f, char *buf, size_t cap)
{
    int32_t len;
    if (rd_exact(f, &len, 4) || len < 0 || (size_t)len >= cap) return -1;
    if (rd_exact(f, buf, (size_t)len)) return -1;
    buf[len] = 0;
    return 0;
}

ORC_API void orc_table_close(orc_table *t)
{
    if (!t) return;
    free(t->cols);
    free(t);
}

/* open_table: creators.jl:7-16 -> read_table_meta + check_column_files */
ORC_API int orc_table_open(const char *path, orc_table **out)
{
    char p[1200];
    snprintf(p, sizeof p, "%s/meta.bin", path);
    FILE *f = fopen(p, "rb");
    if (!f) return fail(ORC_ERR_IO, "Table %s not exists", path);
    orc_table *t = calloc(1, sizeof *t);
    snprintf(t->path, sizeof t->path, "%s", path);
    if (rd_exact(f, &t->format_version, 8) || rd_exact(f, &t->block_size, 8) || rd_exact(f, &t->ncols, 8) ||
        t->ncols < 0 || t->ncols > 100000) {
        fclose(f); orc_table_close(t);
        return fail(ORC_ERR_FORMAT, "bad meta.bin");
    }
    t->cols = calloc((size_t)t->ncols + 1, sizeof(colmeta));
    for (int64_t i = 0; i < t->ncols; i++) {
        colmeta *c = &t->cols[i];
        if (rd_exact(f, &c->id, 8) || rd_string(f, c->name, sizeof c->name) || rd_string(f, c->typestr, sizeof c->typestr)) {
            fclose(f); orc_table_close(t);
            return fail(ORC_ERR_FORMAT, "bad meta.bin");
        }
        int rc = parse_type(c->typestr, strlen(c->typestr), &c->type);
        if (rc) { fclose(f); orc_table_close(t); return rc; }
    }
    fclose(f);
    /* check_column_file: filesystem.jl:56-61 */
    for (int64_t i = 0; i < t->ncols; i++) {
        colmeta *c = &t->cols[i];
        snprintf(p, sizeof p, "%s/%lld.bin", path, (long long)c->id);
        FILE *cf = fopen(p, "rb");
        if (!cf) { orc_table_close(t); return fail(ORC_ERR_IO, "column file '%s' for column %s don't exists", p, c->name); }
        int64_t bs;
        char ts[256];
        if (rd_exact(cf, &bs, 8) || rd_string(cf, ts, sizeof ts)) { fclose(cf); orc_table_close(t); return fail(ORC_ERR_FORMAT, "bad column header %s", p); }
        c->data_start = ftell(cf);
        fclose(cf);
        if (bs != t->block_size) { int64_t tb = t->block_size; orc_table_close(t); return fail(ORC_ERR_FORMAT, "column %s has blocksize %lld, but table has blocksize %lld", c->name, (long long)bs, (long long)tb); }
        if (strcmp(ts, c->typestr) != 0) { orc_table_close(t); return fail(ORC_ERR_FORMAT, "column stored type is %s, but another expected", ts); }
    }
    *out = t;
    return ORC_OK;
}

ORC_API int64_t orc_table_ncols(const orc_table *t) { return t->ncols; }
ORC_API int64_t orc_table_block_size(const orc_table *t) { return t->block_size; }
ORC_API int orc_table_col(const orc_table *t, int64_t i, int64_t *id, char *name, int ncap, char *typestr, int tcap)
{
    if (i < 0 || i >= t->ncols) return fail(ORC_ERR_KEY, "column index out of range");
    *id = t->cols[i].id;
    snprintf(name, (size_t)ncap, "%s", t->cols[i].name);
    snprintf(typestr, (size_t)tcap, "%s", t->cols[i].typestr);
    return ORC_OK;
}
static colmeta *find_col(orc_table *t, int64_t id)
{
    for (int64_t i = 0; i < t->ncols; i++) if (t->cols[i].id == id) return &t->cols[i];
    return NULL;
}

/* ------------------------------------------------------------------------------------------ */
/* block stream: src/io/BlockStreams.jl:68-78 (read_sizes, skip_block), :101-119 (read_block)  */

typedef struct {
    FILE *f;
    uint8_t *comp; size_t comp_cap;
    uint8_t *uncomp; size_t uncomp_cap;
} bstream;

typedef struct { int32_t rows; int64_t origin, compressed; } bsizes;

static int bs_eof(bstream *s)
{
    int c = fgetc(s->f);
    if (c == EOF) return 1;
    ungetc(c, s->f);
    return 0;
}
static int read_sizes(bstream *s, bsizes *z)
{
    if (rd_exact(s->f, &z->rows, 4) || rd_exact(s->f, &z->origin, 8) || rd_exact(s->f, &z->compressed, 8))
        return fail(ORC_ERR_CORRUPT, "truncated block header");
    if (z->rows < 0 || z->origin < 0 || z->compressed < 0 || z->origin > 0x7E000000LL || z->compressed > 0x7F000000LL)
        return fail(ORC_ERR_CORRUPT, "bad block header");
    return ORC_OK;
}
static int skip_block(bstream *s, bsizes *z)
{
    int rc = read_sizes(s, z);
    if (rc) return rc;
    if (fseek(s->f, (long)z->compressed, SEEK_CUR)) return fail(ORC_ERR_IO, "seek failed");
    return ORC_OK;
}
static int read_block(bstream *s, bsizes *z)
{
    int rc = read_sizes(s, z);
    if (rc) return rc;
    if ((size_t)z->compressed > s->comp_cap) { s->comp_cap = (size_t)z->compressed * 2 + 64; s->comp = realloc(s->comp, s->comp_cap); }
    if ((size_t)z->origin + 16 > s->uncomp_cap) { s->uncomp_cap = (size_t)z->origin * 2 + 64; s->uncomp = realloc(s->uncomp, s->uncomp_cap); }
    if (rd_exact(s->f, s->comp, (size_t)z->compressed)) return fail(ORC_ERR_CORRUPT, "truncated block payload");
    int size = orc_lz4_decompress_safe(s->comp, s->uncomp, (int)z->compressed, (int)z->origin);
    if (size != z->origin) return fail(ORC_ERR_CORRUPT, "decompression error");
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* block bodies: src/io/blocks.jl:37-71 ; FlatStringsVectors.jl:61-70                          */

typedef struct {
    coltype t;
    int64_t rows;
    uint8_t *values; size_t values_cap;   /* rows*elsize                               */
    uint8_t *missing; size_t missing_cap; /* rows bytes, 1 = missing (nullable fixed)  */
    int32_t *sizes; size_t str_cap; int64_t *offsets; size_t off_cap;
    uint8_t *chars; size_t chars_cap; int64_t datasize;
} colbuf;

static void *grow(void *p, size_t *cap, size_t need)
{
    if (need <= *cap) return p;
    *cap = need * 2 + 64;
    return realloc(p, *cap);
}

static int read_block_body(const uint8_t *body, int64_t origin, int64_t rows, colbuf *b)
{
    b->rows = rows;
    if (b->t.kind == K_STRING) {
        /* blocks.jl:62-71 : Int32 datasize | rows x Int32 sizes | datasize bytes ; then offsets */
        if (origin < 4 + 4 * rows) return fail(ORC_ERR_CORRUPT, "string body too small");
        int32_t datasize;
        memcpy(&datasize, body, 4);
        if (datasize < 0 || origin != 4 + 4 * rows + datasize) return fail(ORC_ERR_CORRUPT, "string body size mismatch");
        b->sizes = grow(b->sizes, &b->str_cap, (size_t)rows * 4);
        b->offsets = grow(b->offsets, &b->off_cap, (size_t)rows * 8);
        memcpy(b->sizes, body + 4, (size_t)rows * 4);
        b->chars = grow(b->chars, &b->chars_cap, (size_t)datasize + 1);
        memcpy(b->chars, body + 4 + 4 * rows, (size_t)datasize);
        /* unsafe_remake_offsets! FlatStringsVectors.jl:61-70 */
        int64_t off = 0;
        for (int64_t i = 0; i < rows; i++) {
            b->offsets[i] = off;
            if (b->sizes[i] >= 0) off += b->sizes[i];
        }
        if (off != datasize) return fail(ORC_ERR_CORRUPT, "string sizes do not sum to datasize");
        b->datasize = datasize;
        return ORC_OK;
    }
    int es = b->t.elsize;
    if (b->t.nullable) {
        /* blocks.jl:46-60 : BitArray chunks (ceil(rows/64) UInt64, bit=1 => missing) | rows*sizeof(T) */
        int64_t nw = (rows + 63) / 64;
        if (origin != nw * 8 + rows * es) return fail(ORC_ERR_CORRUPT, "missing body size mismatch");
        b->missing = grow(b->missing, &b->missing_cap, (size_t)rows + 1);
        b->values = grow(b->values, &b->values_cap, (size_t)rows * es + 16);
        const uint64_t *w = (const uint64_t *)body;
        for (int64_t i = 0; i < rows; i++) {
            uint64_t word;
            memcpy(&word, &w[i >> 6], 8);
            b->missing[i] = (uint8_t)((word >> (i & 63)) & 1);
        }
        /* Julia's read!(io, ::BitArray) rejects non-zero padding bits in the last chunk */
        if (rows & 63) {
            uint64_t word;
            memcpy(&word, &w[nw - 1], 8);
            if (word >> (rows & 63)) return fail(ORC_ERR_CORRUPT, "BitArray padding bits are not zero");
        }
        memcpy(b->values, body + nw * 8, (size_t)rows * es);
        return ORC_OK;
    }
    /* blocks.jl:37-44 */
    if (origin != rows * es) return fail(ORC_ERR_CORRUPT, "body size mismatch");
    b->values = grow(b->values, &b->values_cap, (size_t)rows * es + 16);
    memcpy(b->values, body, (size_t)rows * es);
    return ORC_OK;
}

static void colbuf_free(colbuf *b)
{
    free(b->values); free(b->missing); free(b->sizes); free(b->offsets); free(b->chars);
    memset(b, 0, sizeof *b);
}

/* ------------------------------------------------------------------------------------------ */
/* plan: serialized SelectionQueue (selection.jl:4-10) + Projection (projection.jl:1-9) whose  */
/* predicate / computed-column nodes are BlockBroadcasting trees (broadcast.jl:6-17) in postfix */

enum { ST_RANGE = 1, ST_INDEXVEC = 2, ST_PRED = 3 };
enum { PJ_COL = 1, PJ_EXPR = 2 };
enum {
    OP_COL = 0x01, OP_I64 = 0x02, OP_F64 = 0x03, OP_STR = 0x04, OP_BOOL = 0x05,
    OP_EQ = 0x10, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE,
    OP_AND = 0x20, OP_OR, OP_XOR, OP_NOT,
    OP_ADD = 0x30, OP_SUB, OP_MUL, OP_DIV, OP_REM, OP_NEG,
    OP_ISMISSING = 0x40, OP_COALESCE,
    OP_STARTSWITH = 0x50, OP_ENDSWITH,
    OP_IN = 0x60
};

typedef struct {
    uint8_t code;
    int64_t i; double f;
    const uint8_t *s; uint32_t slen;
    const int64_t *set; uint32_t nset;
} op_t;

typedef struct { op_t *ops; uint32_t nops; } expr_t;

typedef struct {
    int kind;
    int64_t start, step, stop;          /* RANGE */
    int64_t *idx; int64_t nidx;         /* INDEXVEC (sorted, unique) */
    expr_t e;                           /* PRED */
    /* RangeToProcess state selection.jl:68-75 */
    int64_t offset, first, last;
} stage_t;

typedef struct { int kind; int64_t col; expr_t e; } proj_t;

typedef struct {
    uint32_t nstages; stage_t *stages;
    uint32_t nproj; proj_t *projs;
} plan_t;

typedef struct { const uint8_t *p, *end; int bad; } rdr;
static uint8_t g8(rdr *r) { if (r->p + 1 > r->end) { r->bad = 1; return 0; } return *r->p++; }
static uint32_t g32(rdr *r) { uint32_t v = 0; if (r->p + 4 > r->end) { r->bad = 1; return 0; } memcpy(&v, r->p, 4); r->p += 4; return v; }
static int64_t g64(rdr *r) { int64_t v = 0; if (r->p + 8 > r->end) { r->bad = 1; return 0; } memcpy(&v, r->p, 8); r->p += 8; return v; }
static double gf64(rdr *r) { double v = 0; if (r->p + 8 > r->end) { r->bad = 1; return 0; } memcpy(&v, r->p, 8); r->p += 8; return v; }

static int parse_expr(rdr *r, expr_t *e)
{
    e->nops = g32(r);
    if (r->bad || e->nops > 4096) return fail(ORC_ERR_ARGUMENT, "bad expression");
    e->ops = calloc(e->nops + 1, sizeof(op_t));
    for (uint32_t i = 0; i < e->nops; i++) {
        op_t *o = &e->ops[i];
        o->code = g8(r);
        switch (o->code) {
        case OP_COL: case OP_I64: o->i = g64(r); break;
        case OP_F64: o->f = gf64(r); break;
        case OP_BOOL: o->i = g8(r); break;
        case OP_STR:
            o->slen = g32(r);
            if (r->p + o->slen > r->end) r->bad = 1; else { o->s = r->p; r->p += o->slen; }
            break;
        case OP_IN:
            o->nset = g32(r);
            if (r->p + 8ull * o->nset > r->end) r->bad = 1; else { o->set = (const int64_t *)r->p; r->p += 8ull * o->nset; }
            break;
        default: break;
        }
        if (r->bad) return fail(ORC_ERR_ARGUMENT, "truncated expression");
    }
    return ORC_OK;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : x > y;
}

static void plan_free(plan_t *p)
{
    for (uint32_t i = 0; i < p->nstages; i++) { free(p->stages[i].idx); free(p->stages[i].e.ops); }
    for (uint32_t i = 0; i < p->nproj; i++) free(p->projs[i].e.ops);
    free(p->stages); free(p->projs);
    memset(p, 0, sizeof *p);
}

static int plan_parse(const uint8_t *bytes, int64_t len, plan_t *p)
{
    memset(p, 0, sizeof *p);
    rdr r = { bytes, bytes + len, 0 };
    if (g32(&r) != 0x31504644u) return fail(ORC_ERR_ARGUMENT, "bad plan magic");
    p->nstages = g32(&r);
    if (r.bad || p->nstages > 1024) return fail(ORC_ERR_ARGUMENT, "bad plan");
    p->stages = calloc(p->nstages + 1, sizeof(stage_t));
    for (uint32_t i = 0; i < p->nstages; i++) {
        stage_t *s = &p->stages[i];
        s->kind = g8(&r);
        if (s->kind == ST_RANGE) {
            s->start = g64(&r); s->step = g64(&r); s->stop = g64(&r);
            if (s->step == 0) return fail(ORC_ERR_ARGUMENT, "step cannot be zero");
            int empty = s->step > 0 ? s->start > s->stop : s->start < s->stop;
            /* RangeToProcess ctor: minimum/maximum of an empty range throw (selection.jl:73) */
            if (empty) return fail(ORC_ERR_ARGUMENT, "range must be non-empty");
            s->first = s->start < s->stop ? s->start : s->stop;
            s->last = s->start < s->stop ? s->stop : s->start;
        } else if (s->kind == ST_INDEXVEC) {
            uint32_t n = g32(&r);
            if (r.p + 8ull * n > r.end) return fail(ORC_ERR_ARGUMENT, "truncated plan");
            if (n == 0) return fail(ORC_ERR_ARGUMENT, "reducing over an empty collection is not allowed");
            s->idx = malloc(8ull * n);
            memcpy(s->idx, r.p, 8ull * n);
            r.p += 8ull * n;
            qsort(s->idx, n, 8, cmp_i64);
            int64_t m = 0;
            for (uint32_t k = 0; k < n; k++) if (k == 0 || s->idx[k] != s->idx[m - 1]) s->idx[m++] = s->idx[k];
            s->nidx = m;
            s->first = s->idx[0];
            s->last = s->idx[m - 1];
        } else if (s->kind == ST_PRED) {
            int rc = parse_expr(&r, &s->e);
            if (rc) return rc;
        } else return fail(ORC_ERR_ARGUMENT, "bad stage kind");
        if (r.bad) return fail(ORC_ERR_ARGUMENT, "truncated plan");
    }
    p->nproj = g32(&r);
    if (r.bad || p->nproj > 4096) return fail(ORC_ERR_ARGUMENT, "bad plan");
    p->projs = calloc(p->nproj + 1, sizeof(proj_t));
    for (uint32_t i = 0; i < p->nproj; i++) {
        proj_t *q = &p->projs[i];
        q->kind = g8(&r);
        if (q->kind == PJ_COL) q->col = g64(&r);
        else if (q->kind == PJ_EXPR) { int rc = parse_expr(&r, &q->e); if (rc) return rc; }
        else return fail(ORC_ERR_ARGUMENT, "bad projection kind");
        if (r.bad) return fail(ORC_ERR_ARGUMENT, "truncated plan");
    }
    return ORC_OK;
}

static int stage_member(const stage_t *s, int64_t r)
{
    if (s->kind == ST_RANGE) {
        if (s->step > 0) return r >= s->start && r <= s->stop && (r - s->start) % s->step == 0;
        return r <= s->start && r >= s->stop && (s->start - r) % (-s->step) == 0;
    }
    int64_t lo = 0, hi = s->nidx - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        if (s->idx[mid] == r) return 1;
        if (s->idx[mid] < r) lo = mid + 1; else hi = mid - 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* expression evaluation on the selected rows of one block:                                    */
/*   broadcast.jl:96-118 (_extract_for_eval! gathers), :121-133 (materialize! of the fused tree)*/
/* Values are evaluated column-at-a-time over the n gathered rows.  Semantics are Julia's:     */
/*   exact mixed integer/float comparisons, wrapping integer arithmetic, three-valued logic on  */
/*   missing, `/` always floating, `%` = rem (sign of dividend, DivideError on zero).           */

enum { VT_INT = 1, VT_FLT = 2, VT_BOOL = 3, VT_STR = 4 };

typedef struct {
    int vt;
    int bits, uns;        /* VT_INT: width and signedness ; VT_FLT: bits 32/64 */
    int nullable;
    int is_const;
    int64_t *iv; double *fv; uint8_t *bv; uint8_t *miss;      /* arrays of n (non-const) */
    const uint8_t **sp; int32_t *sl;                          /* strings: pointer + length (-1 missing) */
    int64_t ci; double cf; uint8_t cb; const uint8_t *cs; int32_t csl;
} val_t;

static void val_free(val_t *v)
{
    free(v->iv); free(v->fv); free(v->bv); free(v->miss); free((void *)v->sp); free(v->sl);
    memset(v, 0, sizeof *v);
}

typedef struct {
    orc_table *tbl;
    int ncols;            /* required columns */
    int64_t *col_ids;
    colbuf *bufs;
} blockdata;

static colbuf *bd_col(blockdata *bd, int64_t id)
{
    for (int i = 0; i < bd->ncols; i++) if (bd->col_ids[i] == id) return &bd->bufs[i];
    return NULL;
}

static int64_t wrap_int(int64_t v, int bits, int uns)
{
    if (bits >= 64) return v;
    uint64_t m = (1ull << bits) - 1, u = (uint64_t)v & m;
    if (uns) return (int64_t)u;
    if (u >> (bits - 1)) u |= ~m;
    return (int64_t)u;
}

static int load_col(blockdata *bd, int64_t id, const int64_t *index, int64_t n, val_t *out)
{
    colbuf *b = bd_col(bd, id);
    if (!b) return fail(ORC_ERR_KEY, "column id %lld is not loaded", (long long)id);
    memset(out, 0, sizeof *out);
    out->nullable = b->t.nullable;
    int k = b->t.kind;
    if (k == K_STRING) {
        out->vt = VT_STR;
        out->sp = malloc(sizeof(uint8_t *) * (size_t)(n + 1));
        out->sl = malloc(4 * (size_t)(n + 1));
        if (out->nullable) out->miss = malloc((size_t)n + 1);
        for (int64_t i = 0; i < n; i++) {
            int64_t r = index[i];
            out->sp[i] = b->chars + b->offsets[r];
            out->sl[i] = b->sizes[r];
            if (out->nullable) out->miss[i] = b->sizes[r] < 0;
        }
        return ORC_OK;
    }
    if (out->nullable) {
        out->miss = malloc((size_t)n + 1);
        for (int64_t i = 0; i < n; i++) out->miss[i] = b->missing[index[i]];
    }
    const uint8_t *v = b->values;
#define LOADI(T, BITS, UNS) do { out->vt = VT_INT; out->bits = BITS; out->uns = UNS; out->iv = malloc(8 * (size_t)(n + 1)); \
        for (int64_t i = 0; i < n; i++) { T x; memcpy(&x, v + index[i] * sizeof(T), sizeof(T)); out->iv[i] = (int64_t)x; } } while (0)
    switch (k) {
    case K_I8: LOADI(int8_t, 8, 0); break;
    case K_I16: LOADI(int16_t, 16, 0); break;
    case K_I32: LOADI(int32_t, 32, 0); break;
    case K_I64: case K_DATE: case K_DATETIME: case K_TIME: LOADI(int64_t, 64, 0); break;
    case K_U8: LOADI(uint8_t, 8, 1); break;
    case K_U16: LOADI(uint16_t, 16, 1); break;
    case K_U32: LOADI(uint32_t, 32, 1); break;
    case K_U64: LOADI(uint64_t, 64, 1); break;
    case K_BOOL:
        out->vt = VT_BOOL; out->bv = malloc((size_t)n + 1);
        for (int64_t i = 0; i < n; i++) out->bv[i] = v[index[i]] != 0;
        break;
    case K_F32:
        out->vt = VT_FLT; out->bits = 32; out->fv = malloc(8 * (size_t)(n + 1));
        for (int64_t i = 0; i < n; i++) { float x; memcpy(&x, v + index[i] * 4, 4); out->fv[i] = x; }
        break;
    case K_F64:
        out->vt = VT_FLT; out->bits = 64; out->fv = malloc(8 * (size_t)(n + 1));
        for (int64_t i = 0; i < n; i++) memcpy(&out->fv[i], v + index[i] * 8, 8);
        break;
    default:
        return fail(ORC_ERR_UNSUPPORTED, "column type is not supported in expressions");
    }
#undef LOADI
    if (out->nullable)   /* bytes under a missing bit are garbage (src/common/missings.jl:1): neutralise */
        for (int64_t i = 0; i < n; i++) if (out->miss[i]) { if (out->iv) out->iv[i] = 0; if (out->fv) out->fv[i] = 0; if (out->bv) out->bv[i] = 0; }
    return ORC_OK;
}

/* exact comparison of an int64/uint64 with a double: returns -1,0,1 or 2 for unordered (NaN) */
static int cmp_int_flt(int64_t a, int uns, double b)
{
    if (b != b) return 2;
    if (uns) {
        uint64_t ua = (uint64_t)a;
        if (b < 0) return 1;
        if (b >= 18446744073709551616.0) return -1;
        uint64_t tb = (uint64_t)b;               /* trunc */
        if (ua < tb) return -1;
        if (ua > tb) return 1;
        return (double)tb < b ? -1 : 0;
    }
    if (b >= 9223372036854775808.0) return -1;
    if (b < -9223372036854775808.0) return 1;
    int64_t tb = (int64_t)b;                     /* trunc toward zero, exact in range */
    if (a < tb) return -1;
    if (a > tb) return 1;
    double frac = b - (double)tb;
    return frac > 0 ? -1 : (frac < 0 ? 1 : 0);
}
static int cmp_int_int(int64_t a, int ua, int64_t b, int ub)
{
    if (ua == ub) {
        if (ua) return (uint64_t)a < (uint64_t)b ? -1 : (uint64_t)a > (uint64_t)b;
        return a < b ? -1 : a > b;
    }
    if (ua) { if (b < 0) return 1; return (uint64_t)a < (uint64_t)b ? -1 : (uint64_t)a > (uint64_t)b; }
    if (a < 0) return -1;
    return (uint64_t)a < (uint64_t)b ? -1 : (uint64_t)a > (uint64_t)b;
}
static int cmp_str(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb)
{
    int32_t m = la < lb ? la : lb;
    int c = m ? memcmp(a, b, (size_t)m) : 0;
    if (c) return c < 0 ? -1 : 1;
    return la < lb ? -1 : la > lb;
}

#define VI(v, i) ((v)->is_const ? (v)->ci : (v)->iv[i])
#define VF(v, i) ((v)->is_const ? (v)->cf : (v)->fv[i])
#define VB(v, i) ((v)->is_const ? (v)->cb : (v)->bv[i])
#define VM(v, i) ((v)->nullable && !(v)->is_const ? (v)->miss[i] : 0)

static int apply_cmp(int code, int c)
{
    if (c == 2) return code == OP_NE;   /* NaN: every ordered comparison false, != true */
    switch (code) {
    case OP_EQ: return c == 0;
    case OP_NE: return c != 0;
    case OP_LT: return c < 0;
    case OP_LE: return c <= 0;
    case OP_GT: return c > 0;
    default: return c >= 0;
    }
}

static void result_bool(val_t *r, int64_t n, int nullable)
{
    memset(r, 0, sizeof *r);
    r->vt = VT_BOOL;
    r->nullable = nullable;
    r->bv = calloc((size_t)n + 1, 1);
    if (nullable) r->miss = calloc((size_t)n + 1, 1);
}

static int eval_compare(int code, val_t *a, val_t *b, int64_t n, val_t *r)
{
    int nullable = (a->nullable && !a->is_const) || (b->nullable && !b->is_const);
    if (a->vt == VT_STR || b->vt == VT_STR) {
        if (a->vt != b->vt) {
            /* Julia: ==(::String, ::Number) is false, ordering throws MethodError */
            if (code != OP_EQ && code != OP_NE) return fail(ORC_ERR_UNSUPPORTED, "ordering between String and non-String");
            result_bool(r, n, nullable);
            for (int64_t i = 0; i < n; i++) {
                if (VM(a, i) || VM(b, i)) { r->miss[i] = 1; continue; }
                r->bv[i] = code == OP_NE;
            }
            return ORC_OK;
        }
        result_bool(r, n, nullable);
        for (int64_t i = 0; i < n; i++) {
            if (VM(a, i) || VM(b, i)) { r->miss[i] = 1; continue; }
            const uint8_t *pa = a->is_const ? a->cs : a->sp[i];
            const uint8_t *pb = b->is_const ? b->cs : b->sp[i];
            int32_t la = a->is_const ? a->csl : a->sl[i], lb = b->is_const ? b->csl : b->sl[i];
            r->bv[i] = (uint8_t)apply_cmp(code, cmp_str(pa, la, pb, lb));
        }
        return ORC_OK;
    }
    result_bool(r, n, nullable);
    for (int64_t i = 0; i < n; i++) {
        if (VM(a, i) || VM(b, i)) { r->miss[i] = 1; continue; }
        int c;
        /* Bool participates in comparisons as the integer 0/1 (Julia: Bool <: Integer) */
        int ai = a->vt != VT_FLT, bi = b->vt != VT_FLT;
        int64_t xa = a->vt == VT_BOOL ? VB(a, i) : (a->vt == VT_INT ? VI(a, i) : 0);
        int64_t xb = b->vt == VT_BOOL ? VB(b, i) : (b->vt == VT_INT ? VI(b, i) : 0);
        if (ai && bi) c = cmp_int_int(xa, a->vt == VT_INT && a->uns, xb, b->vt == VT_INT && b->uns);
        else if (ai) c = cmp_int_flt(xa, a->vt == VT_INT && a->uns, VF(b, i));
        else if (bi) { c = cmp_int_flt(xb, b->vt == VT_INT && b->uns, VF(a, i)); if (c != 2) c = -c; }
        else { double x = VF(a, i), y = VF(b, i); c = (x != x || y != y) ? 2 : (x < y ? -1 : x > y); }
        r->bv[i] = (uint8_t)apply_cmp(code, c);
    }
    return ORC_OK;
}

static int eval_logic(int code, val_t *a, val_t *b, int64_t n, val_t *r)
{
    if (a->vt != VT_BOOL || (b && b->vt != VT_BOOL))
        return fail(ORC_ERR_UNSUPPORTED, "bitwise logic is supported on Bool operands only");
    int nullable = (a->nullable && !a->is_const) || (b && b->nullable && !b->is_const);
    result_bool(r, n, nullable);
    for (int64_t i = 0; i < n; i++) {
        int ma = VM(a, i), va = VB(a, i);
        if (code == OP_NOT) { if (ma) r->miss[i] = 1; else r->bv[i] = !va; continue; }
        int mb = VM(b, i), vb = VB(b, i);
        if (code == OP_AND) {
            /* three-valued: false & missing == false */
            if ((!ma && !va) || (!mb && !vb)) r->bv[i] = 0;
            else if (ma || mb) r->miss[i] = 1;
            else r->bv[i] = 1;
        } else if (code == OP_OR) {
            if ((!ma && va) || (!mb && vb)) r->bv[i] = 1;
            else if (ma || mb) r->miss[i] = 1;
            else r->bv[i] = 0;
        } else { /* xor: missing if either missing */
            if (ma || mb) r->miss[i] = 1; else r->bv[i] = (uint8_t)(va ^ vb);
        }
    }
    return ORC_OK;
}

static int eval_arith(int code, val_t *a, val_t *b, int64_t n, val_t *r)
{
    if (a->vt == VT_STR || (b && b->vt == VT_STR)) return fail(ORC_ERR_UNSUPPORTED, "arithmetic on String");
    int nullable = (a->nullable && !a->is_const) || (b && b->nullable && !b->is_const);
    memset(r, 0, sizeof *r);
    r->nullable = nullable;
    if (nullable) r->miss = calloc((size_t)n + 1, 1);
    int af = a->vt == VT_FLT, bf = b && b->vt == VT_FLT;
    /* integer view of Bool operands (Julia promotes Bool to the other integer type, Bool+Bool -> Int64) */
    int abits = a->vt == VT_INT ? a->bits : 0, bbits = b && b->vt == VT_INT ? b->bits : 0;
    int auns = a->vt == VT_INT ? a->uns : 0, buns = b && b->vt == VT_INT ? b->uns : 0;
    if (code == OP_DIV || af || bf) {
        /* floating result: Float64 unless both float operands are <= Float32 / integers with Float32 */
        int bits = 64;
        if (code == OP_DIV && !af && !bf) bits = 64;
        else {
            int fa = af ? a->bits : 0, fb = bf ? b->bits : 0;
            bits = (fa == 64 || fb == 64) ? 64 : 32;
        }
        r->vt = VT_FLT; r->bits = bits;
        r->fv = calloc((size_t)n + 1, 8);
        for (int64_t i = 0; i < n; i++) {
            if (VM(a, i) || (b && VM(b, i))) { r->miss[i] = 1; continue; }
            double x = af ? VF(a, i) : (a->vt == VT_BOOL ? (double)VB(a, i) : (auns ? (double)(uint64_t)VI(a, i) : (double)VI(a, i)));
            double y = 0;
            if (b) y = bf ? VF(b, i) : (b->vt == VT_BOOL ? (double)VB(b, i) : (buns ? (double)(uint64_t)VI(b, i) : (double)VI(b, i)));
            if (bits == 32) { x = (float)x; y = (float)y; }
            double z;
            switch (code) {
            case OP_ADD: z = x + y; break;
            case OP_SUB: z = x - y; break;
            case OP_MUL: z = x * y; break;
            case OP_DIV: z = x / y; break;
            case OP_REM: z = fmod(x, y); break;
            default: z = -x; break;
            }
            if (bits == 32) z = (float)z;
            r->fv[i] = z;
        }
        return ORC_OK;
    }
    /* integer result: promote_type of the operand widths; same width signed+unsigned -> unsigned */
    int bits, uns;
    if (!b) { bits = abits ? abits : 64; uns = auns; }
    else if (abits == 0 && bbits == 0) { bits = 64; uns = 0; }
    else if (abits == 0) { bits = bbits; uns = buns; }
    else if (bbits == 0) { bits = abits; uns = auns; }
    else if (abits == bbits) { bits = abits; uns = auns || buns; }
    else if (abits > bbits) { bits = abits; uns = auns; }
    else { bits = bbits; uns = buns; }
    r->vt = VT_INT; r->bits = bits; r->uns = uns;
    r->iv = calloc((size_t)n + 1, 8);
    for (int64_t i = 0; i < n; i++) {
        if (VM(a, i) || (b && VM(b, i))) { r->miss[i] = 1; continue; }
        int64_t x = a->vt == VT_BOOL ? VB(a, i) : VI(a, i);
        int64_t y = b ? (b->vt == VT_BOOL ? VB(b, i) : VI(b, i)) : 0;
        uint64_t ux = (uint64_t)x, uy = (uint64_t)y, z;
        switch (code) {
        case OP_ADD: z = ux + uy; break;
        case OP_SUB: z = ux - uy; break;
        case OP_MUL: z = ux * uy; break;
        case OP_REM:
            if (y == 0) return fail(ORC_ERR_DIVIDE, "DivideError: integer division error");
            if (uns) z = ux % uy;
            else z = (y == -1) ? 0 : (uint64_t)(x % y);
            break;
        default: z = (uint64_t)0 - ux; break;
        }
        r->iv[i] = wrap_int((int64_t)z, bits, uns);
    }
    return ORC_OK;
}

static int eval_expr(const expr_t *e, blockdata *bd, const int64_t *index, int64_t n, val_t *out)
{
    val_t stack[64];
    int sp = 0, rc = ORC_OK;
    memset(stack, 0, sizeof stack);
    for (uint32_t k = 0; k < e->nops && rc == ORC_OK; k++) {
        const op_t *o = &e->ops[k];
        if (sp >= 60) { rc = fail(ORC_ERR_UNSUPPORTED, "expression too deep"); break; }
        switch (o->code) {
        case OP_COL: rc = load_col(bd, o->i, index, n, &stack[sp]); if (!rc) sp++; break;
        case OP_I64: memset(&stack[sp], 0, sizeof(val_t)); stack[sp].vt = VT_INT; stack[sp].bits = 64; stack[sp].is_const = 1; stack[sp].ci = o->i; sp++; break;
        case OP_F64: memset(&stack[sp], 0, sizeof(val_t)); stack[sp].vt = VT_FLT; stack[sp].bits = 64; stack[sp].is_const = 1; stack[sp].cf = o->f; sp++; break;
        case OP_BOOL: memset(&stack[sp], 0, sizeof(val_t)); stack[sp].vt = VT_BOOL; stack[sp].is_const = 1; stack[sp].cb = (uint8_t)(o->i != 0); sp++; break;
        case OP_STR: memset(&stack[sp], 0, sizeof(val_t)); stack[sp].vt = VT_STR; stack[sp].is_const = 1; stack[sp].cs = o->s; stack[sp].csl = (int32_t)o->slen; sp++; break;
        case OP_EQ: case OP_NE: case OP_LT: case OP_LE: case OP_GT: case OP_GE:
        case OP_AND: case OP_OR: case OP_XOR:
        case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: case OP_REM:
        case OP_STARTSWITH: case OP_ENDSWITH: case OP_COALESCE: {
            if (sp < 2) { rc = fail(ORC_ERR_ARGUMENT, "stack underflow"); break; }
            val_t b = stack[sp - 1], a = stack[sp - 2], r;
            memset(&r, 0, sizeof r);
            if (o->code >= OP_EQ && o->code <= OP_GE) rc = eval_compare(o->code, &a, &b, n, &r);
            else if (o->code >= OP_AND && o->code <= OP_XOR) rc = eval_logic(o->code, &a, &b, n, &r);
            else if (o->code >= OP_ADD && o->code <= OP_REM) rc = eval_arith(o->code, &a, &b, n, &r);
            else if (o->code == OP_COALESCE) {
                /* coalesce(x, default): first non-missing; default must be a constant of x's kind */
                if (!b.is_const || b.vt != a.vt) rc = fail(ORC_ERR_UNSUPPORTED, "coalesce default must be a constant of the same kind");
                else {
                    r = a; memset(&a, 0, sizeof a);      /* steal arrays */
                    if (r.is_const) { /* nothing to do */ }
                    else if (r.nullable) {
                        for (int64_t i = 0; i < n; i++) if (r.miss[i]) {
                            if (r.vt == VT_INT) r.iv[i] = b.ci;
                            else if (r.vt == VT_FLT) r.fv[i] = b.cf;
                            else if (r.vt == VT_BOOL) r.bv[i] = b.cb;
                            else { r.sp[i] = b.cs; r.sl[i] = b.csl; }
                        }
                        free(r.miss); r.miss = NULL; r.nullable = 0;
                    }
                }
            } else {
                /* startswith / endswith (col, const): byte prefix/suffix test */
                if (a.vt != VT_STR || b.vt != VT_STR) rc = fail(ORC_ERR_UNSUPPORTED, "startswith/endswith need String operands");
                else {
                    result_bool(&r, n, (a.nullable && !a.is_const) || (b.nullable && !b.is_const));
                    for (int64_t i = 0; i < n; i++) {
                        if (VM(&a, i) || VM(&b, i)) { r.miss[i] = 1; continue; }
                        const uint8_t *pa = a.is_const ? a.cs : a.sp[i]; int32_t la = a.is_const ? a.csl : a.sl[i];
                        const uint8_t *pb = b.is_const ? b.cs : b.sp[i]; int32_t lb = b.is_const ? b.csl : b.sl[i];
                        if (lb > la) r.bv[i] = 0;
                        else if (o->code == OP_STARTSWITH) r.bv[i] = lb == 0 || memcmp(pa, pb, (size_t)lb) == 0;
                        else r.bv[i] = lb == 0 || memcmp(pa + la - lb, pb, (size_t)lb) == 0;
                    }
                }
            }
            val_free(&a); val_free(&b);
            sp -= 2;
            if (!rc) stack[sp++] = r; else val_free(&r);
            break;
        }
        case OP_NOT: case OP_NEG: case OP_ISMISSING: case OP_IN: {
            if (sp < 1) { rc = fail(ORC_ERR_ARGUMENT, "stack underflow"); break; }
            val_t a = stack[sp - 1], r;
            memset(&r, 0, sizeof r);
            if (o->code == OP_NOT) rc = eval_logic(OP_NOT, &a, NULL, n, &r);
            else if (o->code == OP_NEG) rc = eval_arith(OP_NEG, &a, NULL, n, &r);
            else if (o->code == OP_ISMISSING) {
                result_bool(&r, n, 0);
                for (int64_t i = 0; i < n; i++) r.bv[i] = (uint8_t)VM(&a, i);
            } else {
                if (a.vt != VT_INT && a.vt != VT_BOOL) rc = fail(ORC_ERR_UNSUPPORTED, "in() is supported for integer values");
                else {
                    result_bool(&r, n, a.nullable && !a.is_const);
                    for (int64_t i = 0; i < n; i++) {
                        if (VM(&a, i)) { r.miss[i] = 1; continue; }
                        int64_t x = a.vt == VT_BOOL ? VB(&a, i) : VI(&a, i);
                        int hit = 0;
                        for (uint32_t q = 0; q < o->nset && !hit; q++) {
                            int64_t sv; memcpy(&sv, (const uint8_t *)o->set + 8ull * q, 8);
                            hit = cmp_int_int(x, a.vt == VT_INT && a.uns, sv, 0) == 0;
                        }
                        r.bv[i] = (uint8_t)hit;
                    }
                }
            }
            val_free(&a);
            sp -= 1;
            if (!rc) stack[sp++] = r; else val_free(&r);
            break;
        }
        default: rc = fail(ORC_ERR_UNSUPPORTED, "unknown opcode 0x%02x", o->code); break;
        }
    }
    if (rc == ORC_OK && sp != 1) rc = fail(ORC_ERR_ARGUMENT, "malformed expression");
    if (rc) { for (int i = 0; i < sp; i++) val_free(&stack[i]); return rc; }
    *out = stack[0];
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* required columns: view.jl:183-190 (projection first, then selection), selection.jl:24-35,   */
/* projection.jl:83-97                                                                         */

static void add_unique(int64_t *arr, int *n, int64_t id)
{
    for (int i = 0; i < *n; i++) if (arr[i] == id) return;
    arr[(*n)++] = id;
}
static void expr_cols(const expr_t *e, int64_t *arr, int *n)
{
    for (uint32_t i = 0; i < e->nops; i++) if (e->ops[i].code == OP_COL) add_unique(arr, n, e->ops[i].i);
}

/* ------------------------------------------------------------------------------------------ */
/* BlocksIterator: src/io/blocksiterator.jl:20-44 (DataReader ctor), :46-66 (SizeReader ctor), */
/* :69-78 skipblocks, :80-96 read_cols/skip_cols, :98-121 and :123-145 iterate                  */

typedef struct {
    orc_table *tbl;
    plan_t *plan;
    int size_reader;
    int nreq; int64_t req[256];
    int nsel; int64_t sel[256];
    int nprj; int64_t prj[256];
    bstream streams[256];
    colbuf bufs[256];
    uint8_t *mask; size_t mask_cap;      /* SelectionExecutor.range_buffer selection.jl:87-92 */
    int64_t *index; size_t index_cap;
    int64_t rows;                        /* rows of the current block */
    int64_t nsel_rows;                   /* selected rows in the current block */
    int64_t block_no;                    /* 0-based number of the block just read */
    int64_t next_block;
    int64_t blk_hi;                      /* exclusive upper block bound (block-range driver), <0 = none */
    int closed;
} blkiter;

static int is_finished(plan_t *p)   /* selection.jl:192-196 */
{
    for (uint32_t i = 0; i < p->nstages; i++)
        if (p->stages[i].kind != ST_PRED && p->stages[i].last <= p->stages[i].offset) return 1;
    return 0;
}
static int skip_if_can(plan_t *p, int64_t size_to_skip)   /* selection.jl:177-190 */
{
    if (p->nstages == 0 || p->stages[0].kind == ST_PRED) return 0;
    stage_t *s = &p->stages[0];
    if (s->first - s->offset > size_to_skip) { s->offset += size_to_skip; return 1; }
    return 0;
}
static int isonly_range(plan_t *p)   /* selection.jl:169-175 */
{
    for (uint32_t i = 0; i < p->nstages; i++) if (p->stages[i].kind == ST_PRED) return 0;
    return 1;
}

static void iter_close(blkiter *it)
{
    if (it->closed) return;
    for (int i = 0; i < it->nreq; i++) {
        if (it->streams[i].f) fclose(it->streams[i].f);
        free(it->streams[i].comp); free(it->streams[i].uncomp);
        colbuf_free(&it->bufs[i]);
    }
    free(it->mask); free(it->index);
    it->closed = 1;
}

static int req_index(blkiter *it, int64_t id)
{
    for (int i = 0; i < it->nreq; i++) if (it->req[i] == id) return i;
    return -1;
}

static int iter_open(blkiter *it, orc_table *tbl, plan_t *plan, int size_reader)
{
    memset(it, 0, sizeof *it);
    it->tbl = tbl; it->plan = plan; it->size_reader = size_reader; it->blk_hi = -1;
    int64_t pr[256], sr[256];
    int npr = 0, nsr = 0;
    for (uint32_t i = 0; i < plan->nproj; i++) {
        if (plan->projs[i].kind == PJ_COL) add_unique(pr, &npr, plan->projs[i].col);
        else expr_cols(&plan->projs[i].e, pr, &npr);
        if (npr > 250) return fail(ORC_ERR_UNSUPPORTED, "too many columns");
    }
    for (uint32_t i = 0; i < plan->nstages; i++)
        if (plan->stages[i].kind == ST_PRED) { expr_cols(&plan->stages[i].e, sr, &nsr); if (nsr > 250) return fail(ORC_ERR_UNSUPPORTED, "too many columns"); }
    /* sel_cols / proj_cols split: blocksiterator.jl:27-33 and :47-51 */
    it->nsel = 0;
    for (int i = 0; i < nsr; i++) it->sel[it->nsel++] = sr[i];
    if (it->nsel == 0 && npr > 0) it->sel[it->nsel++] = pr[0];
    it->nprj = 0;
    if (!size_reader)
        for (int i = 0; i < npr; i++) {
            int dup = 0;
            for (int k = 0; k < it->nsel; k++) if (it->sel[k] == pr[i]) dup = 1;
            if (!dup) it->prj[it->nprj++] = pr[i];
        }
    /* required columns: DataReader = unique(proj..., sel...) ; SizeReader = sel_cols only */
    it->nreq = 0;
    if (!size_reader) for (int i = 0; i < npr; i++) add_unique(it->req, &it->nreq, pr[i]);
    for (int i = 0; i < it->nsel; i++) add_unique(it->req, &it->nreq, it->sel[i]);
    for (int i = 0; i < it->nreq; i++) {
        colmeta *c = find_col(tbl, it->req[i]);
        if (!c) { iter_close(it); return fail(ORC_ERR_KEY, "unknown column id %lld", (long long)it->req[i]); }
        char p[1200];
        snprintf(p, sizeof p, "%s/%lld.bin", tbl->path, (long long)c->id);
        it->streams[i].f = fopen(p, "rb");
        if (!it->streams[i].f) { iter_close(it); return fail(ORC_ERR_IO, "cannot open %s", p); }
        /* check_column_head was done at open; position at the first block */
        fseek(it->streams[i].f, (long)c->data_start, SEEK_SET);
        it->bufs[i].t = c->type;
    }
    return ORC_OK;
}

/* apply: selection.jl:161-167 with _apply_to_block :94-111 (range-ish) and :133-157 (predicate) */
static int apply_selection(blkiter *it, int64_t rows, blockdata *bd)
{
    plan_t *p = it->plan;
    it->mask = grow(it->mask, &it->mask_cap, (size_t)rows + 1);
    it->index = grow(it->index, &it->index_cap, ((size_t)rows + 1) * 8);
    memset(it->mask, 1, (size_t)rows);
    for (uint32_t si = 0; si < p->nstages; si++) {
        stage_t *s = &p->stages[si];
        if (s->kind != ST_PRED) {
            int64_t n = 0;
            for (int64_t k = 0; k < rows; k++) if (it->mask[k]) n++;
            int64_t i = 0;
            for (int64_t k = 0; k < rows; k++)
                if (it->mask[k]) { i++; it->mask[k] = (uint8_t)stage_member(s, i + s->offset); }
            s->offset += n;
            if (rows == 0) break;
        } else {
            int64_t n = 0;
            for (int64_t k = 0; k < rows; k++) if (it->mask[k]) it->index[n++] = k;
            val_t r;
            int rc = eval_expr(&s->e, bd, it->index, n, &r);
            if (rc) return rc;
            /* _check_element selection.jl:52-55: eltype must be exactly Bool */
            if (r.vt != VT_BOOL || (r.nullable && !r.is_const)) { val_free(&r); return fail(ORC_ERR_ARGUMENT, "Function for selection must have Bool result type"); }
            for (int64_t i = 0; i < n; i++) it->mask[it->index[i]] = r.is_const ? r.cb : r.bv[i];
            val_free(&r);
            if (rows == 0) break;
        }
    }
    int64_t n = 0;
    for (int64_t k = 0; k < rows; k++) if (it->mask[k]) it->index[n++] = k;
    it->nsel_rows = n;
    return ORC_OK;
}

/* returns 1 when a block with >=1 selected row is ready, 0 at the end, <0 = -error */
static int iter_next(blkiter *it)
{
    plan_t *p = it->plan;
    for (;;) {
        int stop = it->nreq == 0;
        /* skipblocks blocksiterator.jl:69-78 */
        if (!stop) {
            stop = 1;
            while (!bs_eof(&it->streams[0])) {
                if (is_finished(p)) { stop = 1; break; }
                if (!skip_if_can(p, it->tbl->block_size)) { stop = 0; break; }
                for (int i = 0; i < it->nreq; i++) { bsizes z; int rc = skip_block(&it->streams[i], &z); if (rc) return -rc; }
                it->next_block++;
            }
        }
        if (!stop && it->blk_hi >= 0 && it->next_block >= it->blk_hi) stop = 1;
        if (stop) { return 0; }
        it->block_no = it->next_block++;
        bsizes z = {0, 0, 0};
        blockdata bd = { it->tbl, it->nreq, it->req, it->bufs };
        int header_only = it->size_reader && isonly_range(p);
        for (int k = 0; k < it->nsel; k++) {
            int ri = req_index(it, it->sel[k]);
            int rc;
            if (header_only) rc = skip_block(&it->streams[ri], &z);
            else {
                rc = read_block(&it->streams[ri], &z);
                if (!rc) rc = read_block_body(it->streams[ri].uncomp, z.origin, z.rows, &it->bufs[ri]);
            }
            if (rc) return -rc;
        }
        it->rows = z.rows;
        int rc = apply_selection(it, z.rows, &bd);
        if (rc) return -rc;
        if (it->nsel_rows == 0) {
            for (int k = 0; k < it->nprj; k++) { bsizes zz; rc = skip_block(&it->streams[req_index(it, it->prj[k])], &zz); if (rc) return -rc; }
            continue;
        }
        for (int k = 0; k < it->nprj; k++) {
            int ri = req_index(it, it->prj[k]);
            bsizes zz;
            rc = read_block(&it->streams[ri], &zz);
            if (!rc) rc = read_block_body(it->streams[ri].uncomp, zz.origin, zz.rows, &it->bufs[ri]);
            if (rc) return -rc;
            if (zz.rows != z.rows) return -fail(ORC_ERR_CORRUPT, "columns have different block rows");
        }
        return 1;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* consumers                                                                                   */

/* nrow(v): view.jl:192-206 over BlockRowsIterator */
ORC_API int orc_count(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, int64_t *out)
{
    plan_t plan;
    int rc = plan_parse(plan_bytes, plan_len, &plan);
    if (rc) { plan_free(&plan); return rc; }
    blkiter *it = malloc(sizeof *it);
    rc = iter_open(it, tbl, &plan, 1);
    if (rc) { free(it); plan_free(&plan); return rc; }
    int64_t res = 0;
    int st;
    while ((st = iter_next(it)) == 1) res += it->nsel_rows;
    iter_close(it); free(it); plan_free(&plan);
    if (st < 0) return -st;
    *out = res;
    return ORC_OK;
}

/* parity hook: per-row selection flags in table order (1 byte per table row) and the row count */
ORC_API int orc_mask(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, uint8_t *mask, int64_t cap, int64_t *nrows_seen)
{
    plan_t plan;
    int rc = plan_parse(plan_bytes, plan_len, &plan);
    if (rc) { plan_free(&plan); return rc; }
    blkiter *it = malloc(sizeof *it);
    rc = iter_open(it, tbl, &plan, 0);
    if (rc) { free(it); plan_free(&plan); return rc; }
    memset(mask, 0, (size_t)cap);
    int st;
    int64_t maxrow = 0;
    while ((st = iter_next(it)) == 1) {
        int64_t base = it->block_no * tbl->block_size;
        for (int64_t i = 0; i < it->nsel_rows; i++) {
            int64_t g = base + it->index[i];
            if (g >= cap) { iter_close(it); free(it); plan_free(&plan); return fail(ORC_ERR_ARGUMENT, "mask buffer too small"); }
            mask[g] = 1;
        }
        if (base + it->rows > maxrow) maxrow = base + it->rows;
    }
    iter_close(it); free(it); plan_free(&plan);
    if (st < 0) return -st;
    if (nrows_seen) *nrows_seen = maxrow;
    return ORC_OK;
}

/* total rows of the table = sum of block rows of the first column (header walk, misc.jl:6-42) */
ORC_API int orc_table_nrows(orc_table *tbl, int64_t *out)
{
    if (tbl->ncols == 0) { *out = 0; return ORC_OK; }
    char p[1200];
    snprintf(p, sizeof p, "%s/%lld.bin", tbl->path, (long long)tbl->cols[0].id);
    bstream s;
    memset(&s, 0, sizeof s);
    s.f = fopen(p, "rb");
    if (!s.f) return fail(ORC_ERR_IO, "cannot open %s", p);
    fseek(s.f, (long)tbl->cols[0].data_start, SEEK_SET);
    int64_t n = 0;
    while (!bs_eof(&s)) { bsizes z; int rc = skip_block(&s, &z); if (rc) { fclose(s.f); return rc; } n += z.rows; }
    fclose(s.f);
    *out = n;
    return ORC_OK;
}

typedef struct {
    int32_t kind, nullable, elsize, is_expr;
    int64_t nrows;
    uint8_t *values;   /* nrows*elsize (fixed width) */
    uint8_t *missing;  /* nrows bytes, 1 = missing (nullable fixed width) */
    int32_t *sizes;    /* strings: per row byte size, -1 = missing */
    uint8_t *chars;    /* strings: flat bytes */
    int64_t nchars;
    size_t cap_rows, cap_chars;
} orc_col;

typedef struct orc_mat {
    int32_t ncols;
    orc_col *cols;
} orc_mat;

ORC_API void orc_mat_free(orc_mat *m)
{
    if (!m) return;
    for (int i = 0; i < m->ncols; i++) { free(m->cols[i].values); free(m->cols[i].missing); free(m->cols[i].sizes); free(m->cols[i].chars); }
    free(m->cols); free(m);
}
ORC_API int32_t orc_mat_ncols(const orc_mat *m) { return m->ncols; }
ORC_API const orc_col *orc_mat_col(const orc_mat *m, int i) { return &m->cols[i]; }

static void col_reserve(orc_col *c, int64_t more_rows, int64_t more_chars)
{
    size_t need = (size_t)(c->nrows + more_rows);
    if (need > c->cap_rows) {
        c->cap_rows = need * 2 + 64;
        if (c->kind == K_STRING) c->sizes = realloc(c->sizes, c->cap_rows * 4);
        else {
            c->values = realloc(c->values, c->cap_rows * (size_t)c->elsize);
            if (c->nullable) c->missing = realloc(c->missing, c->cap_rows);
        }
    }
    size_t needc = (size_t)(c->nchars + more_chars);
    if (needc > c->cap_chars) { c->cap_chars = needc * 2 + 64; c->chars = realloc(c->chars, c->cap_chars); }
}

/* ColProjExec: projection.jl:130-133 `buffer .= data[name][range]`; strings go through the
 * FlatStringsVector gather FlatStringsVectors.jl:136-157; append! materialization.jl:34-36 */
static int append_col(orc_col *c, colbuf *b, const int64_t *index, int64_t n)
{
    if (b->t.kind == K_STRING) {
        int64_t bytes = 0;
        for (int64_t i = 0; i < n; i++) if (b->sizes[index[i]] > 0) bytes += b->sizes[index[i]];
        col_reserve(c, n, bytes);
        for (int64_t i = 0; i < n; i++) {
            int64_t r = index[i];
            int32_t sz = b->sizes[r];
            c->sizes[c->nrows + i] = sz;
            if (sz > 0) { memcpy(c->chars + c->nchars, b->chars + b->offsets[r], (size_t)sz); c->nchars += sz; }
        }
        c->nrows += n;
        return ORC_OK;
    }
    col_reserve(c, n, 0);
    int es = c->elsize;
    for (int64_t i = 0; i < n; i++) {
        int64_t r = index[i];
        int miss = b->t.nullable ? b->missing[r] : 0;
        if (miss) memset(c->values + (c->nrows + i) * es, 0, (size_t)es);   /* garbage under missing is normalised to 0 */
        else memcpy(c->values + (c->nrows + i) * es, b->values + r * es, (size_t)es);
        if (c->nullable) c->missing[c->nrows + i] = (uint8_t)miss;
    }
    c->nrows += n;
    return ORC_OK;
}

static int append_val(orc_col *c, val_t *v, int64_t n)
{
    if (c->kind == 0) {
        /* first block fixes the result element type of a computed column */
        c->nullable = v->nullable && !v->is_const;
        if (v->vt == VT_INT) {
            static const int sk[] = {K_I8, K_I16, K_I32, K_I64}, uk[] = {K_U8, K_U16, K_U32, K_U64};
            int w = v->bits == 8 ? 0 : v->bits == 16 ? 1 : v->bits == 32 ? 2 : 3;
            c->kind = v->uns ? uk[w] : sk[w]; c->elsize = v->bits / 8;
        } else if (v->vt == VT_FLT) { c->kind = v->bits == 32 ? K_F32 : K_F64; c->elsize = v->bits / 8; }
        else if (v->vt == VT_BOOL) { c->kind = K_BOOL; c->elsize = 1; }
        else { c->kind = K_STRING; c->elsize = 0; }
    }
    if (v->vt == VT_STR) {
        int64_t bytes = 0;
        for (int64_t i = 0; i < n; i++) { int32_t l = v->is_const ? v->csl : v->sl[i]; if (l > 0) bytes += l; }
        col_reserve(c, n, bytes);
        for (int64_t i = 0; i < n; i++) {
            int32_t l = v->is_const ? v->csl : v->sl[i];
            const uint8_t *p = v->is_const ? v->cs : v->sp[i];
            c->sizes[c->nrows + i] = l;
            if (l > 0) { memcpy(c->chars + c->nchars, p, (size_t)l); c->nchars += l; }
        }
        c->nrows += n;
        return ORC_OK;
    }
    col_reserve(c, n, 0);
    for (int64_t i = 0; i < n; i++) {
        uint8_t *dst = c->values + (c->nrows + i) * c->elsize;
        int miss = VM(v, i);
        if (c->nullable) c->missing[c->nrows + i] = (uint8_t)miss;
        if (miss) { memset(dst, 0, (size_t)c->elsize); continue; }
        if (v->vt == VT_INT) { int64_t x = VI(v, i); memcpy(dst, &x, (size_t)c->elsize); }
        else if (v->vt == VT_FLT) { if (c->elsize == 4) { float f = (float)VF(v, i); memcpy(dst, &f, 4); } else { double d = VF(v, i); memcpy(dst, &d, 8); } }
        else dst[0] = VB(v, i);
    }
    c->nrows += n;
    return ORC_OK;
}

/* materialize(v::DFView): materialization.jl:27-40 (pass 1 = nrow for sizehint!, pass 2 = append!);
 * blk_hi >= 0: only blocks [blk_lo, blk_hi) (row-local plans only) -- the thread-per-range driver of the full-size checks */
static int materialize_impl(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, int64_t blk_lo, int64_t blk_hi, orc_mat **out)
{
    plan_t plan;
    int rc = plan_parse(plan_bytes, plan_len, &plan);
    if (rc) { plan_free(&plan); return rc; }
    orc_mat *m = calloc(1, sizeof *m);
    m->ncols = (int32_t)plan.nproj;
    m->cols = calloc(plan.nproj + 1, sizeof(orc_col));
    for (uint32_t i = 0; i < plan.nproj; i++) {
        if (plan.projs[i].kind == PJ_COL) {
            colmeta *c = find_col(tbl, plan.projs[i].col);
            if (!c) { orc_mat_free(m); plan_free(&plan); return fail(ORC_ERR_KEY, "unknown column id"); }
            m->cols[i].kind = c->type.kind; m->cols[i].nullable = c->type.nullable; m->cols[i].elsize = c->type.elsize;
        } else m->cols[i].is_expr = 1;
    }
    blkiter *it = malloc(sizeof *it);
    rc = iter_open(it, tbl, &plan, 0);
    if (rc) { free(it); orc_mat_free(m); plan_free(&plan); return rc; }
    if (blk_hi >= 0) {
        for (uint32_t i = 0; i < plan.nstages && !rc; i++) if (plan.stages[i].kind != ST_PRED) rc = fail(ORC_ERR_UNSUPPORTED, "block ranges need a predicate-only selection");
        for (int64_t b = 0; b < blk_lo && !rc; b++)
            for (int i = 0; i < it->nreq && !rc; i++) { bsizes z; if (bs_eof(&it->streams[i])) break; rc = skip_block(&it->streams[i], &z); }
        if (rc) { iter_close(it); free(it); orc_mat_free(m); plan_free(&plan); return rc; }
        it->next_block = blk_lo;
        it->blk_hi = blk_hi;
    }
    int st;
    while ((st = iter_next(it)) == 1) {
        blockdata bd = { tbl, it->nreq, it->req, it->bufs };
        for (uint32_t i = 0; i < plan.nproj && !rc; i++) {
            if (plan.projs[i].kind == PJ_COL) rc = append_col(&m->cols[i], bd_col(&bd, plan.projs[i].col), it->index, it->nsel_rows);
            else {
                val_t v;
                rc = eval_expr(&plan.projs[i].e, &bd, it->index, it->nsel_rows, &v);
                if (!rc) { rc = append_val(&m->cols[i], &v, it->nsel_rows); val_free(&v); }
            }
        }
        if (rc) break;
    }
    iter_close(it); free(it); plan_free(&plan);
    if (!rc && st < 0) rc = -st;
    if (rc) { orc_mat_free(m); return rc; }
    *out = m;
    return ORC_OK;
}

ORC_API int orc_materialize(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, orc_mat **out)
{
    return materialize_impl(tbl, plan_bytes, plan_len, 0, -1, out);
}

/* ---- content hash of a materialized result, composable over row ranges -------------------------------------------
 * Full-size checks (1e8..1e9 rows) cannot hold both results side by side, so the oracle and the checked result are
 * each reduced to two 64-bit sums per column,  A = sum_i mix(row_i) * (2 i + 1),  B = sum_i mix(row_i)  (mod 2^64),
 * i = position in the result.  A range hashed with local positions re-bases with  A + 2 * base * B,  so a
 * thread-per-block-range driver needs one pass.  mix(row) covers the value bits (0 under a missing flag, as
 * src/common/missings.jl:1 leaves them unspecified), the missing flag, and for strings the size and every byte. */
static inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

ORC_API void orc_hash_fixed(const uint8_t *values, const uint8_t *missing, int elsize, int64_t n, int64_t base, uint64_t *A, uint64_t *B)
{
    uint64_t a = 0, b = 0;
    for (int64_t i = 0; i < n; i++) {
        uint64_t v = 0;
        const int miss = missing ? missing[i] != 0 : 0;
        if (!miss) memcpy(&v, values + i * elsize, (size_t)(elsize > 8 ? 8 : elsize));
        if (!miss && elsize > 8) { uint64_t hi = 0; memcpy(&hi, values + i * elsize + 8, (size_t)(elsize - 8 > 8 ? 8 : elsize - 8)); v ^= mix64(hi); }
        const uint64_t h = mix64(v ^ (miss ? 0xA5A5A5A5A5A5A5A5ull : 0));
        a += h * (2 * (uint64_t)(base + i) + 1);
        b += h;
    }
    *A = a; *B = b;
}

/* `chars` points at the first byte of row 0 of this range */
ORC_API void orc_hash_strings(const int32_t *sizes, const uint8_t *chars, int64_t n, int64_t base, uint64_t *A, uint64_t *B)
{
    uint64_t a = 0, b = 0;
    const uint8_t *p = chars;
    for (int64_t i = 0; i < n; i++) {
        uint64_t f = 0xCBF29CE484222325ull;
        const int32_t sz = sizes[i];
        for (int32_t k = 0; k < sz; k++) f = (f ^ p[k]) * 0x100000001B3ull;
        if (sz > 0) p += sz;
        const uint64_t h = mix64(f ^ ((uint64_t)(uint32_t)sz << 32));
        a += h * (2 * (uint64_t)(base + i) + 1);
        b += h;
    }
    *A = a; *B = b;
}

/* hashes of the rows that blocks [blk_lo, blk_hi) contribute to materialize(plan), positions local to the range */
ORC_API int orc_materialize_hash_blocks(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, int64_t blk_lo, int64_t blk_hi,
                                        uint64_t *A, uint64_t *B, int64_t *nrows, int32_t cap_cols)
{
    orc_mat *m = NULL;
    int rc = materialize_impl(tbl, plan_bytes, plan_len, blk_lo, blk_hi, &m);
    if (rc) return rc;
    if (m->ncols > cap_cols) { orc_mat_free(m); return fail(ORC_ERR_ARGUMENT, "hash buffers too small"); }
    *nrows = m->ncols ? m->cols[0].nrows : 0;
    for (int i = 0; i < m->ncols; i++) {
        orc_col *c = &m->cols[i];
        if (c->kind == K_STRING) orc_hash_strings(c->sizes, c->chars, c->nrows, 0, &A[i], &B[i]);
        else orc_hash_fixed(c->values, c->nullable ? c->missing : NULL, c->elsize, c->nrows, 0, &A[i], &B[i]);
    }
    orc_mat_free(m);
    return ORC_OK;
}

/* aggregates: there is no reduction code in the reference; sum/minimum/maximum/count/mean are
 * Base folds over Base.iterate(::DFColumn) (column.jl:102-126): strict left-to-right over the
 * selected rows in row order.  Both that fold and a Neumaier-compensated sum are reported. */
typedef struct {
    int64_t count;        /* selected rows (including missing)                         */
    int64_t nmissing;     /* selected rows that are missing                            */
    int64_t sum_i64;      /* wrapping two's complement sum (Base.add_sum widens to 64) */
    double sum_fold;      /* left fold in row order                                    */
    double sum_kahan;     /* Neumaier compensated                                      */
    int64_t min_i64, max_i64;
    double min_f64, max_f64;
    int32_t has_nan;
    int32_t kind;         /* K_* of the aggregated values                              */
} orc_agg;

static double jl_min(double a, double b) { if (a != a) return a; if (b != b) return b; if (a < b) return a; if (b < a) return b; return signbit(a) ? a : b; }
static double jl_max(double a, double b) { if (a != a) return a; if (b != b) return b; if (a > b) return a; if (b > a) return b; return signbit(a) ? b : a; }

static int aggregate_impl(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, int32_t proj_idx, int64_t blk_lo, int64_t blk_hi, orc_agg *out)
{
    plan_t plan;
    int rc = plan_parse(plan_bytes, plan_len, &plan);
    if (rc) { plan_free(&plan); return rc; }
    if (proj_idx < 0 || (uint32_t)proj_idx >= plan.nproj) { plan_free(&plan); return fail(ORC_ERR_ARGUMENT, "projection index out of range"); }
    /* DFColumn = one-column view (column.jl:30-37) */
    proj_t keep = plan.projs[proj_idx];
    for (uint32_t i = 0; i < plan.nproj; i++) if ((int32_t)i != proj_idx) free(plan.projs[i].e.ops);
    plan.projs[0] = keep;
    plan.nproj = 1;
    blkiter *it = malloc(sizeof *it);
    rc = iter_open(it, tbl, &plan, 0);
    if (rc) { free(it); plan_free(&plan); return rc; }
    if (blk_hi >= 0) {
        /* block-range driver (multi-threaded CPU baseline): only row-local plans can be split */
        if (!isonly_range(&plan) ? 0 : plan.nstages > 0) rc = fail(ORC_ERR_UNSUPPORTED, "block ranges need a predicate-only selection");
        for (uint32_t i = 0; i < plan.nstages && !rc; i++) if (plan.stages[i].kind != ST_PRED) rc = fail(ORC_ERR_UNSUPPORTED, "block ranges need a predicate-only selection");
        for (int64_t b = 0; b < blk_lo && !rc; b++)
            for (int i = 0; i < it->nreq && !rc; i++) { bsizes z; if (bs_eof(&it->streams[i])) break; rc = skip_block(&it->streams[i], &z); }
        if (rc) { iter_close(it); free(it); plan_free(&plan); return rc; }
        it->next_block = blk_lo;
        it->blk_hi = blk_hi;
    }
    memset(out, 0, sizeof *out);
    double comp = 0;
    int first_i = 1, first_f = 1;
    int st;
    while ((st = iter_next(it)) == 1) {
        blockdata bd = { tbl, it->nreq, it->req, it->bufs };
        val_t v;
        if (keep.kind == PJ_COL) rc = load_col(&bd, keep.col, it->index, it->nsel_rows, &v);
        else rc = eval_expr(&keep.e, &bd, it->index, it->nsel_rows, &v);
        if (rc) break;
        if (v.vt == VT_STR) { val_free(&v); rc = fail(ORC_ERR_UNSUPPORTED, "aggregate over String"); break; }
        for (int64_t i = 0; i < it->nsel_rows; i++) {
            out->count++;
            if (VM(&v, i)) { out->nmissing++; continue; }
            if (v.vt == VT_FLT) {
                double x = VF(&v, i);
                out->kind = v.bits == 32 ? K_F32 : K_F64;
                out->sum_fold += x;
                double t = out->sum_kahan + x;
                if (fabs(out->sum_kahan) >= fabs(x)) comp += (out->sum_kahan - t) + x; else comp += (x - t) + out->sum_kahan;
                out->sum_kahan = t;
                if (x != x) out->has_nan = 1;
                if (first_f) { out->min_f64 = out->max_f64 = x; first_f = 0; }
                else { out->min_f64 = jl_min(out->min_f64, x); out->max_f64 = jl_max(out->max_f64, x); }
            } else {
                int64_t x = v.vt == VT_BOOL ? VB(&v, i) : VI(&v, i);
                int uns = v.vt == VT_INT && v.uns;
                out->kind = v.vt == VT_BOOL ? K_BOOL : (uns ? K_U64 : K_I64);
                out->sum_i64 = (int64_t)((uint64_t)out->sum_i64 + (uint64_t)x);
                if (first_i) { out->min_i64 = out->max_i64 = x; first_i = 0; }
                else if (uns) {
                    if ((uint64_t)x < (uint64_t)out->min_i64) out->min_i64 = x;
                    if ((uint64_t)x > (uint64_t)out->max_i64) out->max_i64 = x;
                } else {
                    if (x < out->min_i64) out->min_i64 = x;
                    if (x > out->max_i64) out->max_i64 = x;
                }
            }
        }
        val_free(&v);
    }
    out->sum_kahan += comp;
    iter_close(it); free(it); plan_free(&plan);
    if (!rc && st < 0) rc = -st;
    return rc;
}

ORC_API int orc_aggregate(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, int32_t proj_idx, orc_agg *out)
{
    return aggregate_impl(tbl, plan_bytes, plan_len, proj_idx, 0, -1, out);
}

/* aggregate over blocks [blk_lo, blk_hi) only -- lets a thread-per-range driver use every host core */
ORC_API int orc_aggregate_blocks(orc_table *tbl, const uint8_t *plan_bytes, int64_t plan_len, int32_t proj_idx,
                                 int64_t blk_lo, int64_t blk_hi, orc_agg *out)
{
    return aggregate_impl(tbl, plan_bytes, plan_len, proj_idx, blk_lo, blk_hi, out);
}

/* codec-level hook: decode one framed block (header + payload) from memory; mirrors
 * read_block BlockStreams.jl:101-119.  Returns the body size or a negative error. */
ORC_API int64_t orc_decode_block(const uint8_t *framed, int64_t len, uint8_t *out, int64_t cap, int32_t *rows)
{
    if (len < 20) return -1;
    int32_t r; int64_t origin, comp;
    memcpy(&r, framed, 4); memcpy(&origin, framed + 4, 8); memcpy(&comp, framed + 12, 8);
    if (comp < 0 || origin < 0 || 20 + comp > len || origin > cap) return -2;
    int size = orc_lz4_decompress_safe(framed + 20, out, (int)comp, (int)origin);
    if (size != origin) return -3;
    *rows = r;
    return origin;
}
