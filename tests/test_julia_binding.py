"""The Julia binding (julia/b200.jl) against the C ABI it binds (include/dfdb_b200.h).

No julia binary exists in the build image, so the shim cannot be executed here; what CAN be checked is that every `ccall`
names a function the header declares, with the right number of arguments, and that every argument / return type has the
width and kind of its C counterpart -- plus the field-by-field layout of the two structs that cross the boundary.  The
ctypes twin (dataframedbs.jl_b200/_capi.py) is held to the same header."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dfdb_b200.h")
SHIM = os.path.join(ROOT, "julia", "b200.jl")


def _strip_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def header_prototypes():
    src = _strip_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"DFDB_API\s+([\w\s\*]+?)\s*\b(dfdb_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argl = [] if args in ("", "void") else [re.sub(r"\s+", " ", a.strip()) for a in args.split(",")]
        protos[name] = (re.sub(r"\s+", " ", ret), argl)
    return protos


def c_kind(decl, is_return=False):
    """(kind, pointee) of a C parameter declaration, name stripped."""
    d = decl.replace("const ", "").strip()
    stars = d.count("*")
    base = d.replace("*", " ").split()
    ty = base[0] if not is_return else " ".join(base)
    if is_return:
        ty = base[0]
    if stars == 0:
        return ("scalar", ty)
    if stars == 1:
        return ("ptr", ty)
    return ("ptrptr", ty)


JL_SCALARS = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64", "uint8_t": "UInt8"}
JL_STRUCTS = {"dfdb_agg": "Agg", "dfdb_outcol": "OutCol"}


def jl_ok(jl, decl, is_return=False):
    kind, ty = c_kind(decl, is_return)
    if kind == "scalar":
        return jl == JL_SCALARS.get(ty)
    if kind == "ptr":
        if ty == "char":
            return jl in ("Cstring", "Ptr{UInt8}")
        if ty in ("void", "dfdb_table", "dfdb_scan"):
            return jl == "Ptr{Cvoid}"
        if ty in JL_STRUCTS:
            return jl in (f"Ref{{{JL_STRUCTS[ty]}}}", f"Ptr{{{JL_STRUCTS[ty]}}}")
        if ty in JL_SCALARS:
            return jl in (f"Ref{{{JL_SCALARS[ty]}}}", f"Ptr{{{JL_SCALARS[ty]}}}")
        return False
    return jl in ("Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}")          # handle / buffer out-parameters


def shim_ccalls():
    src = open(SHIM).read()
    src = "\n".join(line.split("#", 1)[0] if not line.lstrip().startswith("#") else "" for line in src.splitlines())
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\),\s*([\w{}]+),\s*\(([^)]*)\)", src):
        name, ret, args = m.group(1), m.group(2), m.group(3)
        argl = [a.strip() for a in args.split(",") if a.strip()]
        # values passed after the type tuple, up to the matching parenthesis of the ccall
        i, depth, start = m.end(), 1, m.end()
        while depth and i < len(src):
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        vals = src[start:i - 1]
        nvals, d = 0, 0
        cur = ""
        for ch in vals:
            if ch in "([{":
                d += 1
            elif ch in ")]}":
                d -= 1
            if ch == "," and d == 0:
                nvals += bool(cur.strip())
                cur = ""
            else:
                cur += ch
        nvals += bool(cur.strip())
        out.append((name, ret, argl, nvals))
    return out


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = shim_ccalls()
    assert len(calls) >= 15, "the parser lost the shim's ccalls"
    for name, ret, argl, nvals in calls:
        assert name in protos, f"{name} is not declared in include/dfdb_b200.h"
        cret, cargs = protos[name]
        assert len(argl) == len(cargs), f"{name}: {len(argl)} argument types in the ccall, {len(cargs)} in the header"
        assert nvals == len(argl), f"{name}: {nvals} values passed for {len(argl)} argument types"
        assert jl_ok(ret, cret, is_return=True), f"{name}: return type {ret} does not match `{cret}`"
        for i, (j, c) in enumerate(zip(argl, cargs)):
            assert jl_ok(j, c), f"{name}: argument {i + 1} is {j} in the ccall but `{c}` in the header"


def test_the_consumers_the_drop_in_needs_are_bound():
    bound = {c[0] for c in shim_ccalls()}
    need = {"dfdb_init", "dfdb_last_error", "dfdb_table_open", "dfdb_table_close", "dfdb_table_load", "dfdb_table_set_shard",
            "dfdb_scan_prepare", "dfdb_scan_free", "dfdb_scan_count_all", "dfdb_scan_materialize_sizes", "dfdb_scan_materialize",
            "dfdb_scan_aggregate_all", "dfdb_host_alloc", "dfdb_host_free", "dfdb_comm_unique_id", "dfdb_comm_init", "dfdb_comm_destroy"}
    assert need <= bound, f"missing from the shim: {sorted(need - bound)}"


def _header_struct(name):
    src = _strip_comments(open(HEADER).read())
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        ty, rest = stmt.split(None, 1) if "*" not in stmt.split()[0] else (stmt.split()[0], stmt.split(None, 1)[1])
        for nm in rest.split(","):
            nm = nm.strip()
            ptr = nm.startswith("*") or ty.endswith("*")
            fields.append((nm.lstrip("*"), (ty.rstrip("*") + " *") if ptr else ty))
    return fields


def _shim_struct(name):
    src = open(SHIM).read()
    body = re.search(r"struct %s\n(.*?)\nend" % name, src, flags=re.S).group(1)
    fields = []
    for part in re.split(r"[;\n]", body):
        part = part.strip()
        if part:
            nm, ty = part.split("::")
            fields.append((nm.strip(), ty.strip()))
    return fields


def test_struct_layouts_agree_field_by_field():
    from dfdb_b200 import _capi
    for cname, jname, ctype in (("dfdb_agg", "Agg", _capi.Agg), ("dfdb_outcol", "OutCol", _capi.OutCol)):
        hf, jf = _header_struct(cname), _shim_struct(jname)
        assert [f[0] for f in hf] == [f[0] for f in jf] == [f[0] for f in ctype._fields_], cname
        for (nm, cty), (_, jty), (_, pty) in zip(hf, jf, ctype._fields_):
            if cty.endswith("*"):
                assert jty.startswith("Ptr{") and C.sizeof(pty) == 8, (cname, nm)
            else:
                assert jty == JL_SCALARS[cty], (cname, nm, jty, cty)
                assert C.sizeof(pty) == {"Int32": 4, "Int64": 8, "Float64": 8, "UInt8": 1}[jty], (cname, nm)
    assert C.sizeof(_capi.Agg) == 80 or C.sizeof(_capi.Agg) == 88          # 9 x 8 + 2 x 4 (+ padding)


def test_ctypes_twin_declares_every_header_symbol():
    from dfdb_b200 import _capi
    protos = header_prototypes()
    assert set(protos) == set(_capi.SYMBOLS), sorted(set(protos) ^ set(_capi.SYMBOLS))
    for name, (cret, cargs) in protos.items():
        res, args = _capi.SYMBOLS[name]
        assert len(args) == len(cargs), name
