"""julia/FVMCuda.jl cannot be executed here (no Julia in the image), so its foreign calls are checked statically:
every `ccall((:name, LIB), Ret, (Args...), ...)` must name a function include/fvmcuda.h declares, with the same
arity, the same C argument types and the same return type; and the binding must reach every setter of the ABI
(a drop-in that silently skips e.g. the source term would compute S = 0 without an error)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_TYPES = {
    "fvm_handle": "ptr_void", "fvm_wire_handle": "ptr_void", "void*": "ptr_void", "const void*": "ptr_void",
    "fvm_handle*": "ptr_ptr_void", "fvm_wire_handle*": "ptr_ptr_void", "void**": "ptr_ptr_void",
    "double*": "ptr_f64", "const double*": "ptr_f64", "int32_t*": "ptr_i32", "const int32_t*": "ptr_i32",
    "int64_t*": "ptr_i64", "const int64_t*": "ptr_i64", "uint8_t*": "ptr_u8", "const uint8_t*": "ptr_u8",
    "char*": "cstring", "const char*": "cstring", "int32_t": "i32", "int64_t": "i64", "double": "f64", "uint32_t": "u32", "void": "void",
}
JL_TYPES = {"Ptr{Cvoid}": "ptr_void", "Ptr{Ptr{Cvoid}}": "ptr_ptr_void", "Ptr{Float64}": "ptr_f64", "Ptr{Int32}": "ptr_i32",
            "Ptr{Int64}": "ptr_i64", "Ptr{UInt8}": "ptr_u8", "Cstring": "cstring", "Int32": "i32", "Int64": "i64", "Float64": "f64",
            "UInt32": "u32", "Cvoid": "void"}


def header_prototypes():
    src = open(os.path.join(ROOT, "include", "fvmcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int32_t|uint32_t|const char\s*\*|void)\s+(fvm_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        params = []
        args = " ".join(args.split())
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"^(.*?[\s\*])(\w+)$", a)  # strip the parameter name
                typ = (m.group(1) if m else a).strip()
                typ = re.sub(r"\s*\*", "*", typ)
                params.append(C_TYPES[typ])
        protos[name] = (C_TYPES[re.sub(r"\s*\*", "*", ret.strip())], params)
    return protos


def julia_ccalls():
    src = open(os.path.join(ROOT, "julia", "FVMCuda.jl")).read()
    src = re.sub(r"#=.*?=#", " ", src, flags=re.S)
    calls = []
    for name, ret, args in re.findall(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(([^)]*)\)", src, flags=re.S):
        params = [JL_TYPES[a.strip()] for a in args.split(",") if a.strip()]
        calls.append((name, JL_TYPES[ret], params))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(protos) >= 60 and len(calls) >= 35
    for name, ret, params in calls:
        assert name in protos, "%s is not declared in include/fvmcuda.h" % name
        cret, cparams = protos[name]
        assert ret == cret, "%s: return type %s vs %s in the header" % (name, ret, cret)
        assert len(params) == len(cparams), "%s: %d arguments vs %d in the header" % (name, len(params), len(cparams))
        for k, (j, c) in enumerate(zip(params, cparams)):
            # a typed Julia array may be handed to a `void*` parameter (fvm_host_register, fvm_wire_put)
            assert j == c or (c == "ptr_void" and j.startswith("ptr_")), "%s: argument %d is %s, the header says %s" % (name, k, j, c)


def test_binding_reaches_every_setter_and_the_hot_path():
    protos = header_prototypes()
    used = {name for name, _, _ in julia_ccalls()}
    setters = {n for n in protos if n.startswith("fvm_set_")} - {"fvm_set_ghost_nodes", "fvm_set_halo", "fvm_set_profiling"}  # sharding / profiling: Python-side only
    assert setters <= used, "setters never called from julia/FVMCuda.jl: %s" % sorted(setters - used)
    hot = {"fvm_create", "fvm_finalize", "fvm_destroy", "fvm_rhs", "fvm_apply_dirichlet", "fvm_assemble", "fvm_get_csr", "fvm_spmv",
           "fvm_tsit5", "fvm_tsit5_adaptive", "fvm_krylov", "fvm_jacobian", "fvm_get_jacobian_csr", "fvm_last_error"}
    assert hot <= used, sorted(hot - used)


def test_binding_covers_the_registry_of_the_python_mirror():
    """the Julia functor structs mirror finitevolumemethod.jl_b200/functors.py one to one (same names)"""
    import fvm_b200 as G
    src = open(os.path.join(ROOT, "julia", "FVMCuda.jl")).read()
    names = ["ConstantDiffusion", "TabulatedDiffusion", "PowerDiffusion", "AdvectionDiffusionFlux", "KellerSegelFlux", "ZeroSource",
             "LinearSource", "LogisticSource", "TabulatedSource", "GrayScottSource", "BrusselatorSource", "KellerSegelSource", "Const",
             "AffineU", "ExpSaturation", "LinearXY", "ExpXYT"]
    for n in names:
        assert hasattr(G, n) and re.search(r"struct %s\b" % n, src), n
    assert "FVMSystem" in src and "subproblems(prob::FVMSystem)" in src
