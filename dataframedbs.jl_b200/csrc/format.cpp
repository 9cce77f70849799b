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

// ---- block index + zone map sidecar: <table>/<id>.zmap --------------------------------------------------------------
// The reference's format has no index (skip_block walks the headers, BlockStreams.jl:74-78) and no per-block statistics.
// The sidecar is OPTIONAL and lives beside the column file, which stays byte-identical, so the reference still opens the
// table (it only ever opens meta.bin and <id>.bin).  It is trusted only while the column file has the size and the
// modification time recorded in it; otherwise it is ignored and the headers are walked as before.
//   char[8] "DFDBZM01" | i64 block_size | i64 nblocks | i64 col_id | i64 bin_size | i64 bin_mtime_ns | i32 kind, nullable, elsize, cls
//   nblocks x { i64 file_off | i32 rows | i32 flags | i64 origin | i64 compressed | i64 null_count | u64 min_bits | u64 max_bits }
namespace {
struct __attribute__((packed)) ZmHeader { char magic[8]; int64_t block_size, nblocks, col_id, bin_size, bin_mtime_ns; int32_t kind, nullable, elsize, cls; };
struct __attribute__((packed)) ZmEntry { int64_t file_off; int32_t rows, flags; int64_t origin, compressed, null_count; uint64_t min_bits, max_bits; };

int64_t mtime_ns(const struct stat &st) { return (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec; }

// fills c.blocks (+ totals) and c.zones from a valid sidecar; false = no usable sidecar
bool zonemap_read(const std::string &base, Column &c, int64_t block_size, const struct stat &bin)
{
    File f;
    if (!f.open_ro(base + "/" + std::to_string(c.id) + ".zmap")) return false;
    ZmHeader h;
    if (!f.read(&h, sizeof h) || memcmp(h.magic, "DFDBZM01", 8) != 0) return false;
    if (h.block_size != block_size || h.col_id != c.id || h.bin_size != (int64_t)bin.st_size || h.bin_mtime_ns != mtime_ns(bin)) return false;
    if (h.kind != c.type.kind || h.nullable != (c.type.nullable ? 1 : 0) || h.elsize != c.type.elsize) return false;
    if (h.nblocks < 0 || h.nblocks > (1 << 28) || f.size != (int64_t)(sizeof h + (size_t)h.nblocks * sizeof(ZmEntry))) return false;
    std::vector<ZmEntry> e((size_t)h.nblocks);
    if (h.nblocks && !f.read(e.data(), e.size() * sizeof(ZmEntry))) return false;
    // the index must tile the column file exactly: data_start | 20-byte header | payload | 20-byte header | payload ...
    int64_t pos = c.data_start;
    for (const ZmEntry &z : e) {
        if (z.rows < 0 || z.origin < 0 || z.compressed < 0 || z.origin > 0x7E000000LL || z.file_off != pos + 20) return false;
        pos = z.file_off + z.compressed;
    }
    if (pos != (int64_t)bin.st_size) return false;
    c.blocks.clear();
    c.zones.clear();
    c.total_compressed = c.total_origin = 0;
    for (const ZmEntry &z : e) {
        c.blocks.push_back(BlockInfo{z.file_off, z.rows, z.origin, z.compressed});
        c.total_compressed += z.compressed;
        c.total_origin += z.origin;
        ZoneEntry ze;
        ze.min_bits = z.min_bits; ze.max_bits = z.max_bits; ze.null_count = z.null_count; ze.flags = z.flags;
        c.zones.push_back(ze);
    }
    c.index_from_sidecar = true;
    return true;
}
}  // namespace

int zonemap_write(const dfdb_table *t, const Column &c)
{
    if (c.zones.size() != c.blocks.size()) return fail(DFDB_ERR_STATE, "column %s has no zone maps to write", c.name.c_str());
    const std::string bin = t->path + "/" + std::to_string(c.id) + ".bin", dst = t->path + "/" + std::to_string(c.id) + ".zmap", tmp = dst + ".tmp";
    struct stat st;
    if (stat(bin.c_str(), &st)) return fail(DFDB_ERR_IO, "cannot stat %s", bin.c_str());
    ZmHeader h;
    memcpy(h.magic, "DFDBZM01", 8);
    h.block_size = t->block_size; h.nblocks = (int64_t)c.blocks.size(); h.col_id = c.id; h.bin_size = (int64_t)st.st_size; h.bin_mtime_ns = mtime_ns(st);
    h.kind = c.type.kind; h.nullable = c.type.nullable ? 1 : 0; h.elsize = c.type.elsize; h.cls = value_class(c.type.kind);
    std::vector<ZmEntry> e(c.blocks.size());
    for (size_t b = 0; b < e.size(); b++) {
        e[b].file_off = c.blocks[b].file_off; e[b].rows = c.blocks[b].rows; e[b].flags = c.zones[b].flags; e[b].origin = c.blocks[b].origin;
        e[b].compressed = c.blocks[b].compressed; e[b].null_count = c.zones[b].null_count; e[b].min_bits = c.zones[b].min_bits; e[b].max_bits = c.zones[b].max_bits;
    }
    FILE *fp = fopen(tmp.c_str(), "wb");
    if (!fp) return fail(DFDB_ERR_IO, "cannot write %s", tmp.c_str());
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1 && (e.empty() || fwrite(e.data(), sizeof(ZmEntry), e.size(), fp) == e.size());
    ok = fclose(fp) == 0 && ok;
    if (!ok || rename(tmp.c_str(), dst.c_str())) { unlink(tmp.c_str()); return fail(DFDB_ERR_IO, "cannot write %s", dst.c_str()); }
    return DFDB_OK;
}

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
        {
            struct stat bst;
            if (fstat(f.fd, &bst) == 0 && zonemap_read(base, c, t->block_size, bst)) {
                for (const BlockInfo &bi : c.blocks) rows_total += bi.rows;
                f.pos = f.size;                           // the sidecar's index replaces the header walk
            }
        }
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
