"""The Julia binding (julia/B200ExaModels.jl) cannot be executed here (no Julia), so it is checked
statically against include/iexa.h: operator codes, the ccall'ed symbols and their argument counts."""
import os
import re

import iexa_b200 as ex
from conftest import ROOT

JL = open(os.path.join(ROOT, "julia", "B200ExaModels.jl")).read()
HDR = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "iexa.h")).read(), flags=re.S)


def test_operator_codes_match_the_header():
    enum = dict((k, int(v)) for k, v in re.findall(r"IEXA_OP_([A-Z0-9]+)\s*=\s*(\d+)", HDR))
    jl = dict((k, int(v)) for k, v in re.findall(r":([A-Za-z0-9+\-*/^]+)\s*=>\s*(\d+)", JL))
    sym = {"+": "ADD", "-": "SUB", "*": "MUL", "/": "DIV", "^": "POW", "neg": "NEG", "pos": "POS"}
    for name, code in jl.items():
        hname = sym.get(name, name.upper())
        if name == "csch":                                   # the reference's :csch => csc quirk (operators.jl:41)
            assert code == enum["CSC"]
            continue
        assert enum[hname] == code, (name, code, enum.get(hname))
    # every operator of src/operators.jl:2-46 is bound
    assert set(ex.expr._OP_MAPPINGS) <= set(jl)
    # and the Python table agrees with the header too
    for k, v in ex.OP.items():
        assert enum[k] == v


def test_every_ccall_names_a_declared_symbol_with_the_right_arity():
    decl = {}
    for ret, name, args in re.findall(r"(\w[\w\s\*]*?)\b(iexa_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", HDR):
        args = args.strip()
        decl[name] = 0 if args in ("", "void") else len(args.split(","))
    calls = re.findall(r"ccall\(\(\s*:(iexa_[a-z0-9_]+)\s*,\s*LIB\)\s*,\s*\w+\s*,\s*\(([^)]*)\)", JL)
    calls += [(m.group(1), m.group(2)) for m in re.finditer(
        r"\(\$\(QuoteNode\(sym\)\), LIB\), Int32, \(([^)]*)\)", JL) for _ in ()]
    assert len(calls) >= 15
    for name, argtypes in calls:
        assert name in decl, f"{name} is not declared in iexa.h"
        n = len([a for a in argtypes.split(",") if a.strip()])
        assert n == decl[name], f"{name}: ccall passes {n} arguments, the header declares {decl[name]}"


def test_every_entry_point_of_the_header_is_bound_or_deliberately_left_out():
    decl = set(re.findall(r"\b(iexa_[a-z0-9_]+)\s*\(", HDR))
    bound = set(re.findall(r":(iexa_[a-z0-9_]+)", JL))
    not_needed = {"iexa_version", "iexa_debug_cache_stats", "iexa_debug_get_column", "iexa_debug_codegen_source_of", "iexa_debug_codegen_compile_of", "iexa_engine_note", "iexa_algorithmic_bytes", "iexa_launches_per_call",   # reporting only
                  "iexa_debug_codegen_compile", "iexa_debug_codegen_source", "iexa_debug_set_class_mode",  # debug
                  "iexa_set_vector",      # x0 / y0 are ordinary Julia vectors of the NLPModelMeta
                  "iexa_csr_create"}      # iexa_csr_create_keyed with keys = NULL
    assert decl - bound <= not_needed, sorted(decl - bound - not_needed)


def test_struct_mirrors():
    assert re.search(r"struct IexaNode.*?op::Int32\s+a::Int32\s+b::Int32\s+pad::Int32\s+c::Float64", JL, re.S)
    assert re.search(r"struct IexaIndex.*?base::Int64\s+nterms::Int32\s+col::NTuple\{4,Int32\}\s+pad::Int32\s+coef::NTuple\{4,Int64\}", JL, re.S)
    fields = re.search(r"struct IexaMeta(.*?)end", JL, re.S).group(1)
    assert len(re.findall(r"::Int64", fields)) == 11 and len(re.findall(r"::Int32", fields)) == 6
