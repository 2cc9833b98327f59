"""Structural check of julia/VoronoiFVMB200.jl (Julia itself is not installed in this image): every `ccall` names an entry point
declared in include/vfvm_b200.h and exported by libvfvmb200.so, with the declared number of arguments and matching C types; the shim
defines methods with the reference's real signatures for the three dispatch points of the Newton path; blocks are balanced."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "julia", "VoronoiFVMB200.jl")
HEADER = os.path.join(ROOT, "include", "vfvm_b200.h")
LIB = os.path.join(ROOT, "voronoifvm.jl_b200", "libvfvmb200.so")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def header_decls():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|void|const char\s*\*)\s+(vfvm_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        params = [] if args in ("", "void") else _split_top(args)
        decls[name] = (ret.replace(" ", ""), params)
    return decls


def c_kind(param):
    p = param.strip()
    if "*" in p or "[" in p:
        return "ptr"
    if re.match(r"(const\s+)?double\b", p):
        return "double"
    if re.match(r"(const\s+)?int64_t\b", p):
        return "int64"
    if re.match(r"(const\s+)?int\b", p):
        return "int"
    raise AssertionError(f"unclassified C parameter {param!r}")


def julia_kind(t):
    t = t.strip()
    if t.startswith("Ptr{") or t.startswith("Ref{") or t == "Cstring":
        return "ptr"
    return {"Cint": "int", "Cdouble": "double", "Float64": "double", "Int64": "int64", "Clonglong": "int64"}[t]


def shim_ccalls():
    txt = open(SHIM).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(", txt):
        name, ret = m.group(1), m.group(2)
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[i], 0)
            i += 1
        types = _split_top(txt[m.end() : i - 1])
        j, depth = i, 1  # the remaining call arguments up to the ccall's closing parenthesis
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[j], 0)
            j += 1
        values = _split_top(txt[i:j - 1].lstrip(","))
        calls.append((name, ret, types, values))
    return calls


def test_every_ccall_matches_the_header_and_the_library():
    decls = header_decls()
    lib = ctypes.CDLL(LIB)
    calls = shim_ccalls()
    assert len(calls) >= 25
    for name, ret, types, values in calls:
        assert name in decls, f"{name} is not declared in include/vfvm_b200.h"
        assert hasattr(lib, name), f"{name} is not exported by libvfvmb200.so"
        cret, params = decls[name]
        assert len(types) == len(params), f"{name}: shim passes {len(types)} argument types, header declares {len(params)}"
        assert len(values) == len(types), f"{name}: {len(values)} values for {len(types)} argument types"
        for t, p in zip(types, params):
            assert julia_kind(t) == c_kind(p), f"{name}: Julia type {t} does not match C parameter {p!r}"
        assert {"int": "Cint", "void": "Cvoid", "constchar*": "Cstring"}[cret] == ret, f"{name}: return type {ret} vs {cret}"


def test_shim_binds_the_whole_newton_path():
    names = {c[0] for c in shim_ccalls()}
    needed = {"vfvm_create", "vfvm_destroy", "vfvm_set_grid", "vfvm_build_geometry", "vfvm_set_system", "vfvm_set_physics", "vfvm_set_nodal_source", "vfvm_set_legacy_bc",
              "vfvm_set_bc_entries", "vfvm_build_pattern", "vfvm_set_vector", "vfvm_get_vector", "vfvm_copy_vector", "vfvm_init_dirichlet", "vfvm_assemble",
              "vfvm_eval_res_jac", "vfvm_linsolve_setup", "vfvm_linsolve", "vfvm_linsolve_status", "vfvm_amg_set_options", "vfvm_newton_update", "vfvm_vector_norms",
              "vfvm_timings", "vfvm_pattern_size", "vfvm_get_pattern_csc", "vfvm_get_nzval_csc", "vfvm_last_error"}
    assert needed <= names, sorted(needed - names)


def test_shim_defines_the_reference_dispatch_points_with_their_real_signatures():
    txt = " ".join(open(SHIM).read().split())
    # src/vfvm_assembly.jl:520-534
    assert re.search(r"function VoronoiFVM\.eval_and_assemble\( system, U::AbstractMatrix\{Tv\}, UOld::AbstractMatrix\{Tv\}, F::AbstractMatrix\{Tv\}, matrix::B200Matrix, "
                     r"generic_matrix::Union\{AbstractMatrix, Nothing\}, dudp, time, tstep, λ, data, params::AbstractVector; edge_cutoff = 0\.0, \) where \{Tv\}", txt)
    # src/vfvm_linsolve.jl:6
    assert "function VoronoiFVM._solve_linear!(u, state, nlhistory, control, method_linear, A::B200Matrix, b, reuse_precs)" in txt
    # src/vfvm_solver.jl:13-23
    assert re.search(r"function VoronoiFVM\.solve_step!\( state::B200State, solution, oldsol, control, time, tstep, embedparam, params, istep_factorize \)", txt)
    # src/vfvm_state.jl:99-103 and src/vfvm_solver.jl:665-668
    assert "function VoronoiFVM.SystemState(backend::B200, system::VoronoiFVM.AbstractSystem;" in txt
    assert "function CommonSolve.solve(system::VoronoiFVM.AbstractSystem, backend::B200;" in txt


def test_every_registered_physics_id_has_a_julia_struct():
    hdr = open(HEADER).read()
    shim = open(SHIM).read()
    ids = {}
    for m in re.finditer(r"#define VFVM_(FLUX|REACTION|STORAGE|SOURCE|BREACTION|EDGEREACTION|BSTORAGE)_(\w+) (\d+)", hdr):
        ids[(m.group(1), int(m.group(3)))] = m.group(2)
    have = {}
    for m in re.finditer(r"struct (\w+)(?:\{[^}]*\})? <: Registered(Flux|Reaction|Storage|Source|BReaction|EdgeReaction|BStorage)\b", shim):
        name, kind = m.group(1), m.group(2).upper()
        pid = re.search(rf"physics_id\(::{name}\) = (\d+)", shim)
        if pid:
            have[(kind, int(pid.group(1)))] = name
    missing = {k: v for k, v in ids.items() if k not in have}
    assert not missing, f"registered ids without a Julia struct: {missing}"


def test_blocks_are_balanced():
    txt = open(SHIM).read()
    txt = re.sub(r'"""(.|\n)*?"""', '""', txt)
    txt = re.sub(r'"(\\.|[^"\\\n])*"', '""', txt)
    txt = re.sub(r"#.*", "", txt)
    for a, b in ("()", "[]", "{}"):
        assert txt.count(a) == txt.count(b), f"unbalanced {a}{b}"
    # [begin:end] / a[end] indexing is not used in the shim, so every `end` closes a block
    opens = len(re.findall(r"(?<![\w.!])(?:module|function|struct|if|for|while|begin|let|try|do|quote)(?![\w!])", txt))
    opens += len(re.findall(r"abstract type", txt))  # `abstract type X end`
    depth, comp_for = 0, 0  # comprehensions: a `for` inside square brackets has no `end`
    for m in re.finditer(r"\[|\]|(?<![\w.!])for(?![\w!])", txt):
        if m.group(0) == "[":
            depth += 1
        elif m.group(0) == "]":
            depth -= 1
        elif depth > 0:
            comp_for += 1
    opens -= comp_for
    ends = len(re.findall(r"(?<![\w.!])end(?![\w!])", txt))
    assert opens == ends, f"{opens} block openers vs {ends} `end`"
