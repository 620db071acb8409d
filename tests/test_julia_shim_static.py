"""The Julia shim (julia/CIAOAlgorithmsCUDA.jl) cannot be executed in this image (no Julia), so its `ccall`s are checked
statically against the prototypes of include/ciao_cuda.h: every symbol it binds is declared, with the same number of
arguments and C-compatible argument and return types.  Catches the shim drifting from the ABI."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Julia ccall type  →  the C parameter types it may be bound to
JL2C = {
    "Ptr{Cvoid}": {"ciao_ctx *", "void *", "const void *", "ciao_perturb_fn"},   # the last one: a @cfunction pointer
    "Ref{Ptr{Cvoid}}": {"ciao_ctx **"},
    "Cint": {"int"},
    "Int64": {"int64_t"},
    "UInt64": {"uint64_t"},
    "Float64": {"double"},
    "Ptr{Float64}": {"double *", "const double *"},
    "Ref{Float64}": {"double *"},
    "Ptr{Int64}": {"int64_t *", "const int64_t *"},
    "Ref{Int64}": {"int64_t *"},
    "Ptr{Cint}": {"int *"},
    "Ref{Cint}": {"int *"},
}
RET = {"Cint": "int", "Cstring": "const char *"}


def split_top(s):
    """split on commas that are not inside braces/parentheses"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def header_prototypes():
    src = open(os.path.join(ROOT, "include", "ciao_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\n\s*((?:const\s+)?[A-Za-z_0-9]+\s*\**)\s*(ciao_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        ptypes = []
        for p in split_top(params):
            p = " ".join(p.split())
            if p in ("void", ""):
                continue
            t = re.sub(r"\s*[A-Za-z_0-9]+$", "", p) if not p.endswith("*") else p       # drop the parameter name
            t = re.sub(r"\s*\*", " *", t).replace("* *", "**").strip()
            ptypes.append(t)
        protos[name] = (" ".join(ret.split()).replace(" *", " *"), ptypes)
    return protos


def shim_ccalls():
    src = open(os.path.join(ROOT, "julia", "CIAOAlgorithmsCUDA.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(:(ciao_[a-z_0-9]+),\s*libciao\),\s*([A-Za-z]+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:                                   # the argument-type tuple, balanced
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tuple_txt = src[m.end():i - 1]
        j, depth = i, 1                                # the rest of the ccall: the argument values
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        values = split_top(src[i:j - 1].lstrip(", \n"))
        calls.append((m.group(1), m.group(2), [t for t in split_top(tuple_txt) if t], values))
    return calls


def test_every_ccall_matches_a_header_prototype():
    protos, calls = header_prototypes(), shim_ccalls()
    assert len(protos) >= 30 and len(calls) >= 20
    for name, ret, jl_types, values in calls:
        assert name in protos, f"{name} is not declared in include/ciao_cuda.h"
        c_ret, c_types = protos[name]
        assert RET[ret] == c_ret, f"{name}: returns {c_ret}, the shim says {ret}"
        assert len(jl_types) == len(c_types), f"{name}: {len(c_types)} parameters in the header, {len(jl_types)} types in the ccall"
        assert len(values) == len(jl_types), f"{name}: {len(jl_types)} argument types but {len(values)} values"
        for k, (jt, ct) in enumerate(zip(jl_types, c_types)):
            assert jt in JL2C, f"{name}: unknown Julia type {jt}"
            assert ct in JL2C[jt], f"{name}: argument {k + 1} is `{ct}` in the header, `{jt}` in the ccall"


def test_the_shim_binds_every_solver_entry_point():
    bound = {c[0] for c in shim_ccalls()}
    for sym in ("ciao_create", "ciao_destroy", "ciao_last_error", "ciao_set_rows", "ciao_set_blocks", "ciao_set_reg", "ciao_get_vec",
                "ciao_svrg_init", "ciao_svrg_epoch", "ciao_saga_init", "ciao_saga_steps", "ciao_finito_init", "ciao_finito_steps",
                "ciao_lfinito_init", "ciao_lfinito_outer", "ciao_proshi_init", "ciao_proshi_steps", "ciao_proshi_solution",
                "ciao_finito_adaptive_init_cb", "ciao_finito_adaptive_steps", "ciao_finito_adaptive_get"):
        assert sym in bound, sym


def _strip_strings_and_comments(src):
    out, i, n = [], 0, len(src)
    while i < n:
        ch = src[i]
        if ch == "#":                                   # line comment (the shim has no #= =# blocks)
            while i < n and src[i] != "\n":
                i += 1
        elif ch == '"':                                 # string literal, escapes skipped; $(...) interpolation treated as text
            i += 1
            while i < n and src[i] != '"':
                i += 2 if src[i] == "\\" else 1
            i += 1
            out.append('""')
        elif ch == "'" and i + 2 < n and src[i + 2] == "'":   # character literal
            i += 3
            out.append("' '")
        else:
            out.append(ch)
            i += 1
    return "".join(out)


def test_blocks_and_brackets_balance():
    """No Julia here to parse the shim, so at least: brackets nest properly and every block opener outside brackets
    (function, if, for, while, struct, begin, module, let, try, do, quote, macro) has its `end`."""
    src = _strip_strings_and_comments(open(os.path.join(ROOT, "julia", "CIAOAlgorithmsCUDA.jl")).read())
    pairs = {")": "(", "]": "[", "}": "{"}
    stack, blocks = [], []
    openers = {"function", "if", "for", "while", "struct", "begin", "module", "let", "try", "do", "quote", "macro"}
    word = re.compile("[A-Za-z_\\u0080-\\uffff][A-Za-z_0-9!\\u0080-\\uffff]*|[()\\[\\]{}]")
    for lineno, line in enumerate(src.splitlines(), 1):
        for m in word.finditer(line):
            tok = m.group(0)
            if tok in "([{":
                stack.append((tok, lineno))
            elif tok in ")]}":
                assert stack and stack[-1][0] == pairs[tok], f"line {lineno}: unmatched {tok}"
                stack.pop()
            elif not stack:                              # keywords inside brackets are comprehensions / generators / a[end]
                prev = line[:m.start()].rstrip()
                if prev.endswith(".") or prev.endswith(":"):     # field access x.begin / symbol :for — not a keyword
                    continue
                if tok in openers:
                    blocks.append((tok, lineno))
                elif tok == "end":
                    assert blocks, f"line {lineno}: `end` without an open block"
                    blocks.pop()
    assert not stack, f"unclosed bracket opened at line {stack[-1][1]}"
    assert not blocks, f"unclosed `{blocks[-1][0]}` opened at line {blocks[-1][1]}"
