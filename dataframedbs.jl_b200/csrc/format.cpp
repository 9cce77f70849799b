// format.cpp -- host-side reader of the reference's on-disk table format and block indexer.
//
//   <table>/meta.bin   read_table_meta      /root/reference/src/io/table_io.jl:21-33
//   strings            read_string          /root/reference/src/io/common_io.jl:5-8  (Int32 nbytes + bytes)
//   <table>/<id>.bin   check_column_head    /root/reference/src/io/filesystem.jl:47-54 (Int64 block_size + typestring)
//   block headers      read_sizes           /root/reference/src/io/BlockStreams.jl:68-72 (Int32 rows, Int64 origin, Int64 compressed)
//   typestrings        parse_typestring     /root/reference/src/columntypes/base.jl:41-74, complex.jl:1-20
//
// The format has no footer or index (locating block k means walking k headers, BlockStreams.jl:74-78),
// so the walk is done once here and kept as a per-column block index.
#include <cstdarg>
#include <cstdlib>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "internal.hpp"

namespace dfdb {

static thread_local char g_err[1024];

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
const char *last_error() { return g_err; }

static const struct { const char *name; int kind, size; } k_prims[] = {
    {"Int8", DFDB_I8, 1}, {"Int16", DFDB_I16, 2}, {"Int32", DFDB_I32, 4}, {"Int64", DFDB_I64, 8}, {"Int128", DFDB_I128, 16},
    {"UInt8", DFDB_U8, 1}, {"UInt16", DFDB_U16, 2}, {"UInt32", DFDB_U32, 4}, {"UInt64", DFDB_U64, 8}, {"UInt128", DFDB_U128, 16},
    {"Float16", DFDB_F16, 2}, {"Float32", DFDB_F32, 4}, {"Float64", DFDB_F64, 8}, {"Bool", DFDB_BOOL, 1}, {"Char", DFDB_CHAR, 4},
    {"String", DFDB_STRING, 0}, {"Date", DFDB_DATE, 8}, {"DateTime", DFDB_DATETIME, 8}, {"Time", DFDB_TIME, 8},
};

int value_class(int kind)
{
    switch (kind) {
    case DFDB_I8: case DFDB_I16: case DFDB_I32: case DFDB_I64: case DFDB_DATE: case DFDB_DATETIME: case DFDB_TIME: return VC_INT;
    case DFDB_U8: case DFDB_U16: case DFDB_U32: case DFDB_U64: return VC_UINT;
    case DFDB_F32: case DFDB_F64: return VC_FLT;
    case DFDB_BOOL: return VC_BOOL;
    case DFDB_STRING: return VC_STR;
    default: return VC_NONE;
    }
}

int parse_typestring(const char *s, size_t n, ColType *out)
{
    while (n && (*s == ' ' || *s == '\t')) { s++; n--; }
    while (n && (s[n - 1] == ' ' || s[n - 1] == '\t')) n--;
    if (n == 0 || s[0] == '(') return fail(DFDB_ERR_FORMAT, "typename parse error");
    const char *brace = static_cast<const char *>(memchr(s, '(', n));
    if (!brace) {
        for (const auto &p : k_prims)
            if (strlen(p.name) == n && memcmp(p.name, s, n) == 0) {
                out->kind = p.kind;
                out->nullable = false;
                out->elsize = p.size;
                out->align = p.size > 8 ? 16 : (p.size ? p.size : 1);
                return DFDB_OK;
            }
        return fail(DFDB_ERR_FORMAT, "Undefined column type: %.*s", (int)n, s);
    }
    if (s[n - 1] != ')') return fail(DFDB_ERR_FORMAT, "typename parse error");
    size_t hn = (size_t)(brace - s);
    const char *inner = brace + 1;
    size_t in = n - hn - 2;
    if (hn == 7 && memcmp(s, "Missing", 7) == 0) {
        int rc = parse_typestring(inner, in, out);
        if (rc) return rc;
        if (out->nullable) return fail(DFDB_ERR_FORMAT, "nested Missing type");
        out->nullable = true;
        return DFDB_OK;
    }
    if (hn == 5 && memcmp(s, "Tuple", 5) == 0) {
        // Julia lays out isbits tuples like C structs: natural alignment, size padded to the max alignment
        int depth = 0, size = 0, maxal = 1, count = 0;
        size_t start = 0;
        for (size_t i = 0; i <= in; i++) {
            char c = i < in ? inner[i] : ',';
            if (c == '(') depth++;
            else if (c == ')') depth--;
            else if (c == ',' && depth == 0) {
                if (i > start) {
                    ColType e;
                    int rc = parse_typestring(inner + start, i - start, &e);
                    if (rc) return rc;
                    if (e.kind == DFDB_STRING || e.nullable) return fail(DFDB_ERR_FORMAT, "Unsupported tuple element type");
                    size = (size + e.align - 1) / e.align * e.align + e.elsize;
                    if (e.align > maxal) maxal = e.align;
                    count++;
                }
                start = i + 1;
            }
        }
        if (!count) return fail(DFDB_ERR_FORMAT, "Undefined column type: Tuple");
        out->kind = DFDB_TUPLE;
        out->nullable = false;
        out->align = maxal;
        out->elsize = (size + maxal - 1) / maxal * maxal;
        return DFDB_OK;
    }
    return fail(DFDB_ERR_FORMAT, "Undefined column type: %.*s", (int)n, s);
}

namespace {
struct File {
    int fd = -1;
    int64_t size = 0, pos = 0;
    ~File() { if (fd >= 0) close(fd); }
    bool open_ro(const std::string &p)
    {
        fd = ::open(p.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st)) return false;
        size = st.st_size;
        return true;
    }
    bool read(void *dst, size_t n)
    {
        size_t got = 0;
        while (got < n) {
            ssize_t r = pread(fd, static_cast<char *>(dst) + got, n - got, pos + (int64_t)got);
            if (r <= 0) return false;
            got += (size_t)r;
        }
        pos += (int64_t)n;
        return true;
    }
    bool read_string(std::string *out)
    {
        int32_t len;
        if (!read(&len, 4) || len < 0 || len > (1 << 20)) return false;
        out->resize((size_t)len);
        return len == 0 || read(&(*out)[0], (size_t)len);
    }
};
}  // namespace

int table_open_host(const char *path, dfdb_table **out)
{
    struct stat st;
    if (stat(path, &st) || !S_ISDIR(st.st_mode)) return fail(DFDB_ERR_IO, "Table %s don't exists", path);
    std::string base(path);
    while (base.size() > 1 && base.back() == '/') base.pop_back();
    File mf;
    if (!mf.open_ro(base + "/meta.bin")) return fail(DFDB_ERR_IO, "Meta file %s/meta.bin don't exists", base.c_str());
    auto *t = new dfdb_table();
    t->path = base;
    int64_t ncols = 0;
    if (!mf.read(&t->format_version, 8) || !mf.read(&t->block_size, 8) || !mf.read(&ncols, 8) || ncols < 0 || ncols > 100000 ||
        t->block_size <= 0 || t->block_size > (1 << 28)) {
        delete t;
        return fail(DFDB_ERR_FORMAT, "bad meta.bin in %s", base.c_str());
    }
    t->cols.resize((size_t)ncols);
    for (auto &c : t->cols) {
        if (!mf.read(&c.id, 8) || !mf.read_string(&c.name) || !mf.read_string(&c.typestr)) {
            delete t;
            return fail(DFDB_ERR_FORMAT, "bad meta.bin in %s", base.c_str());
        }
        int rc = parse_typestring(c.typestr.data(), c.typestr.size(), &c.type);
        if (rc) { delete t; return rc; }
    }
    // check_column_file + block index
    bool first = true;
    for (auto &c : t->cols) {
        std::string p = base + "/" + std::to_string(c.id) + ".bin";
        File f;
        if (!f.open_ro(p)) {
            std::string nm = c.name;
            delete t;
            return fail(DFDB_ERR_IO, "column file '%s' for column %s don't exists", p.c_str(), nm.c_str());
        }
        int64_t bs;
        std::string ts;
        if (!f.read(&bs, 8) || !f.read_string(&ts)) { delete t; return fail(DFDB_ERR_FORMAT, "bad column header in %s", p.c_str()); }
        if (bs != t->block_size) {
            int rc = fail(DFDB_ERR_FORMAT, "column %s has blocksize %lld, but table has blocksize %lld", c.name.c_str(), (long long)bs,
                          (long long)t->block_size);
            delete t;
            return rc;
        }
        if (ts != c.typestr) {
            int rc = fail(DFDB_ERR_FORMAT, "column %s stored type is %s, but %s expected", c.name.c_str(), ts.c_str(), c.typestr.c_str());
            delete t;
            return rc;
        }
        c.data_start = f.pos;
        int64_t rows_total = 0;
        while (f.pos < f.size) {
            struct __attribute__((packed)) { int32_t rows; int64_t origin, compressed; } h;
            if (!f.read(&h, 20)) { delete t; return fail(DFDB_ERR_CORRUPT, "truncated block header in %s", p.c_str()); }
            if (h.rows < 0 || h.origin < 0 || h.compressed < 0 || h.origin > 0x7E000000LL || f.pos + h.compressed > f.size) {
                delete t;
                return fail(DFDB_ERR_CORRUPT, "bad block header in %s at byte %lld", p.c_str(), (long long)(f.pos - 20));
            }
            c.blocks.push_back(BlockInfo{f.pos, h.rows, h.origin, h.compressed});
            c.total_compressed += h.compressed;
            c.total_origin += h.origin;
            rows_total += h.rows;
            f.pos += h.compressed;
        }
        if (first) {
            t->nrows = rows_total;
            t->nblocks = (int64_t)c.blocks.size();
            first = false;
        } else if (rows_total != t->nrows || (int64_t)c.blocks.size() != t->nblocks) {
            int rc = fail(DFDB_ERR_CORRUPT, "column %s has %lld rows in %zu blocks, table has %lld rows in %lld blocks", c.name.c_str(),
                          (long long)rows_total, c.blocks.size(), (long long)t->nrows, (long long)t->nblocks);
            delete t;
            return rc;
        }
        // every block but the last must hold block_size rows (columns.jl:130-181 write path)
        for (size_t b = 0; b + 1 < c.blocks.size(); b++)
            if (c.blocks[b].rows != t->block_size) {
                int rc = fail(DFDB_ERR_CORRUPT, "column %s block %zu has %d rows, expected %lld", c.name.c_str(), b, c.blocks[b].rows,
                              (long long)t->block_size);
                delete t;
                return rc;
            }
    }
    t->blk_lo = 0;
    t->blk_hi = t->nblocks;
    *out = t;
    return DFDB_OK;
}

}  // namespace dfdb
