"""The Fortran shims cannot be compiled here (no Fortran compiler in the image or on the GPU box), so they are checked statically against
the C ABI they bind: every `bind(C, name=...)` interface must name a function that include/*.h declares and the library exports, with the
same number of arguments; every call of such an interface inside a shim procedure must pass that many arguments; every subroutine /
function / module / interface block must be closed by its own `end`."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _join_continuations(text):
    out, cur = [], ""
    for line in text.splitlines():
        code = line.split("!")[0].rstrip() if '"' not in line else re.sub(r'!(?=[^"]*$).*', "", line).rstrip()
        if not code.strip():
            continue
        s = code.strip()
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1] + " "
        else:
            out.append(cur + s); cur = ""
    assert cur == ""
    return out


def _split_args(s):
    args, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            args.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        args.append(cur.strip())
    return args


def _c_prototypes():
    protos = {}
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"\b(?:int|long|void|const char \*)\s*\*?\s*(ab200_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
            a = m.group(2).strip()
            protos[m.group(1)] = 0 if a in ("", "void") else len(_split_args(a))
    return protos


def test_shims_match_the_c_abi(built):
    protos = _c_prototypes()
    assert len(protos) > 60
    lib = ctypes.CDLL(os.path.join(ROOT, "aeroapps_b200", "libaeroapps_b200.so"))
    nbind = ncall = 0
    for f in sorted(glob.glob(os.path.join(ROOT, "aeroapps_b200", "fortran", "*.F90"))):
        lines = _join_continuations(open(f).read())
        iface = {}
        for ln in lines:
            m = re.match(r"(?i)\s*(?:integer\(c_int\)|integer\(c_long\)|type\(c_ptr\))?\s*(function|subroutine)\s+(\w+)\s*\((.*)\)\s*bind\(C,\s*name=\"(\w+)\"\)", ln)
            if m:
                name, args, cname = m.group(2), _split_args(m.group(3)), m.group(4)
                assert cname in protos, f"{os.path.basename(f)}: {cname} is not declared in include/*.h"
                assert hasattr(lib, cname), f"{cname} is not exported by the library"
                assert len(args) == protos[cname], f"{os.path.basename(f)}: interface {name} has {len(args)} arguments, the C prototype {protos[cname]}"
                iface[name.lower()] = len(args); nbind += 1
        for ln in lines:
            for m in re.finditer(r"(?i)\b(ab200_\w+)\s*\(", ln):
                name = m.group(1).lower()
                if name not in iface or re.search(r"(?i)\b(function|subroutine)\s+" + name, ln):
                    continue
                depth, j = 0, m.end() - 1
                for k in range(j, len(ln)):
                    depth += ln[k] == "("; depth -= ln[k] == ")"
                    if depth == 0:
                        break
                n = len(_split_args(ln[j + 1:k]))
                assert n == iface[name], f"{os.path.basename(f)}: call of {name} passes {n} arguments, its interface has {iface[name]}: {ln[:120]}"
                ncall += 1
        # block structure
        stack = []
        for ln in lines:
            s = ln.strip().lower()
            m = re.match(r"end\s*(subroutine|function|module|interface)\b\s*(\w*)", s)
            if m:
                kind, nm = stack.pop()
                assert kind == m.group(1), f"{os.path.basename(f)}: 'end {m.group(1)}' closes a {kind} ({nm})"
                if m.group(2) and kind != "interface":
                    assert m.group(2) == nm, f"{os.path.basename(f)}: end {kind} {m.group(2)} closes {nm}"
                continue
            m = re.match(r"(?:(?:integer\(c_int\)|integer\(c_long\)|type\(c_ptr\)|logical|integer|real\*8|real\(c_double\))\s+)?(subroutine|function)\s+(\w+)", s)
            if m and "module procedure" not in s:
                stack.append((m.group(1), m.group(2))); continue
            m = re.match(r"module\s+(\w+)$", s)
            if m and not s.startswith("module procedure"):
                stack.append(("module", m.group(1))); continue
            if re.match(r"interface\b", s):
                stack.append(("interface", s[9:].strip()))
        assert not stack, f"{os.path.basename(f)}: unclosed {stack}"
    assert nbind >= 45 and ncall >= 45, (nbind, ncall)


def test_shim_lines_fit_the_free_form_limit():
    """free-form Fortran lines are limited to 132 characters unless a compiler flag lifts the limit; AeroApps' rtms components build without one"""
    for f in glob.glob(os.path.join(ROOT, "aeroapps_b200", "fortran", "*.F90")):
        for k, line in enumerate(open(f).read().splitlines(), 1):
            assert len(line) <= 132, f"{os.path.basename(f)}:{k} has {len(line)} characters"
