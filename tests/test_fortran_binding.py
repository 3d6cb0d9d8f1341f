"""fortran/rays_b200_m.f90 against include/rays_b200.h, without a Fortran compiler (there is none in this image):

* every `type, bind(C)` is parsed and compared FIELD BY FIELD (kind, array length, order) with the C struct of the same name;
* every `bind(C, name=...)` interface is compared with the C prototype (argument count, by-value vs by-reference, kinds);
* every routine the replacement `trace_rays` calls is defined in the file or imported from a RAYS module with `use`;
* the module variables it imports exist in the reference modules (when /root/reference is present, i.e. in the build container).
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F90 = os.path.join(ROOT, "fortran", "rays_b200_m.f90")
HDR = os.path.join(ROOT, "include", "rays_b200.h")
REF = "/root/reference/RAYS_project"

C2F = {"int32_t": "integer(c_int32_t)", "int64_t": "integer(c_int64_t)", "double": "real(c_double)", "int": "integer(c_int)",
       "size_t": "integer(c_size_t)"}


def f90_source():
    """free-form source with continuation lines joined and comments stripped"""
    out, cur = [], ""
    for ln in open(F90):
        ln = ln.split("!")[0].rstrip() if "'" not in ln.split("!")[0] or ln.count("'") % 2 == 0 else ln.rstrip()
        if not ln.strip():
            continue
        s = ln.strip()
        if s.startswith("&"):
            s = s[1:].lstrip()
        if cur:
            cur += " " + s
        else:
            cur = s
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        out.append(cur)
        cur = ""
    return out


def c_structs():
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    macros = {m.group(1): m.group(2) for m in re.finditer(r"#define\s+(\w+)\s+\(?([^\n)]+)\)?", txt)}

    def ev(e):
        e = e.strip()
        for _ in range(4):
            for k, v in macros.items():
                e = re.sub(r"\b%s\b" % k, "(" + v + ")", e)
        return int(eval(e))
    structs = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", txt, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ptr = "*" in decl
            decl_nc = decl.replace("const ", "")
            typ, rest = decl_nc.split(" ", 1)
            if typ == "char" and ptr:
                typ = "double"      # any pointer -> type(c_ptr) below
            for name in rest.split(","):
                name = name.strip()
                isptr = ptr and ("*" in name or "*" in typ or decl_nc.count("*") > 0 and name == rest.split(",")[0].strip())
                name_clean = name.replace("*", "").strip()
                dim = 1
                am = re.match(r"(\w+)\[(.+)\]$", name_clean)
                if am:
                    name_clean, dim = am.group(1), ev(am.group(2))
                if "*" in name or (ptr and "*" in rest and name == rest.split(",")[0].strip() and "*" in rest.split(",")[0]):
                    ftype = "type(c_ptr)"
                elif typ in C2F:
                    ftype = C2F[typ]
                else:
                    ftype = f"type({typ})"
                fields.append((name_clean.lower(), ftype, dim))
        structs[m.group(1)] = fields
    return structs


def f_types():
    types, cur, name = {}, None, None
    for ln in f90_source():
        m = re.match(r"type\s*,\s*bind\(C\)\s*::\s*(\w+)", ln, flags=re.I)
        if m:
            name, cur = m.group(1), []
            continue
        if cur is not None and re.match(r"end type", ln, flags=re.I):
            types[name] = cur
            cur = None
            continue
        if cur is not None:
            typ, names = ln.split("::")
            typ = typ.strip().replace(" ", "")
            for nm in re.findall(r"(\w+(?:\([^)]*\))?)", names):
                am = re.match(r"(\w+)\((\w+)\)$", nm)
                dim = 1
                if am:
                    nm, d = am.group(1), am.group(2)
                    dim = {"RAYS_NSPECIES": 6}.get(d, int(d) if d.isdigit() else None)
                cur.append((nm.lower(), typ, dim))
    return types


def test_bind_c_types_match_the_header_field_by_field():
    cs, fs = c_structs(), f_types()
    assert len(fs) >= 12
    for name, ff in fs.items():
        assert name in cs, f"Fortran type {name} has no C struct"
        cf = cs[name]
        assert [x[0] for x in ff] == [x[0] for x in cf], f"{name}: field names/order differ:\n F {[x[0] for x in ff]}\n C {[x[0] for x in cf]}"
        for (fn, ft, fd), (cn, ct, cd) in zip(ff, cf):
            assert ft.lower() == ct.lower(), f"{name}.{fn}: Fortran {ft} vs C {ct}"
            assert fd == cd, f"{name}.{fn}: array length {fd} vs {cd}"
    # the structs the replacement trace_rays fills must all be mirrored
    for must in ("rays_cfg", "rays_fan", "rays_results", "rays_slab_eq", "rays_solovev_eq", "rays_axisym_eq", "rays_mirror_eq", "rays_spline1d",
                 "rays_spline2d", "rays_deposition"):
        assert must in fs


def c_prototypes():
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*(?:const\s+)?(\w+)\s*(\*?)\s*(rays_b200_\w+)\(([^;]*?)\);", txt, flags=re.S | re.M):
        args = " ".join(m.group(4).split())
        alist = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
        protos[m.group(3)] = (m.group(1) + m.group(2), alist)
    return protos


def test_interfaces_match_the_c_prototypes():
    protos = c_prototypes()
    src = f90_source()
    n_checked = 0
    i = 0
    while i < len(src):
        m = re.match(r"(.*)function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)", src[i], flags=re.I)
        if not m:
            i += 1
            continue
        ret, fname, fargs, cname = m.group(1).strip(), m.group(2), [a.strip() for a in m.group(3).split(",") if a.strip()], m.group(4)
        assert fname == cname and cname in protos, cname
        cret, cargs = protos[cname]
        assert len(fargs) == len(cargs), f"{cname}: {len(fargs)} Fortran arguments vs {len(cargs)} in C"
        decl = {}
        i += 1
        while not re.match(r"end function", src[i], flags=re.I):
            if "::" in src[i] and not src[i].lower().startswith("import"):
                typ, names = src[i].split("::")
                for nm in names.split(","):
                    decl[nm.strip().split("(")[0]] = typ.replace(" ", "").lower()
            i += 1
        for fa, ca in zip(fargs, cargs):
            t = decl[fa]
            by_value = "value" in t
            c_is_ptr = "*" in ca
            if t.startswith("type(c_ptr)"):      # type(c_ptr), value <-> T *;  type(c_ptr) by reference <-> T **
                assert ca.count("*") == (1 if by_value else 2), f"{cname}({fa}): {t} vs C `{ca}`"
                continue
            assert by_value != c_is_ptr, f"{cname}({fa}): by value = {by_value} but C argument is `{ca}`"
            ctype = ca.replace("const ", "").replace("*", " ").split()[0]
            if ctype in C2F and not (c_is_ptr and ctype in ("void", "char")):
                assert t.startswith(C2F[ctype].replace(" ", "").lower()), f"{cname}({fa}): {t} vs C {ca}"
            elif ctype.startswith("rays_"):
                assert t.startswith(f"type({ctype})"), f"{cname}({fa}): {t} vs C {ca}"
        if cret == "int":
            assert ret.replace(" ", "").lower() == "integer(c_int)"
        n_checked += 1
    assert n_checked >= 10


def test_every_routine_the_replacement_calls_is_defined():
    src = f90_source()
    text = "\n".join(src)
    defined = {m.group(1).lower() for m in re.finditer(r"^\s*(?:integer\s+function|function|subroutine|integer\(c_int\)\s+function|type\(c_ptr\)\s+function)\s+(\w+)", text, flags=re.I | re.M)}
    imported = set()
    for ln in src:
        m = re.match(r"use\s+(\w+)\s*,\s*only\s*:\s*(.*)", ln, flags=re.I)
        if m:
            for it in m.group(2).split(","):
                imported.add(it.split("=>")[0].strip().lower())
    called = {m.group(1).lower() for m in re.finditer(r"\bcall\s+(\w+)", text, flags=re.I)}
    intrinsic = {"c_f_pointer", "get_environment_variable"}
    missing = sorted(c for c in called if c not in defined and c not in imported and c not in intrinsic)
    assert not missing, f"called but neither defined in the file nor imported: {missing}"
    for must in ("pack_slab_eq", "pack_solovev_eq", "pack_axisym_toroid_eq", "pack_multiple_mirror_eq", "allocate_ray_results"):
        assert must in defined


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_imported_module_variables_exist_in_the_reference():
    """every `use <module>, only : a, b => c` names a public entity of that reference module"""
    src = f90_source()
    n = 0
    for ln in src:
        m = re.match(r"use\s+(\w+)\s*,\s*only\s*:\s*(.*)", ln, flags=re.I)
        if not m or m.group(1).lower() in ("iso_c_binding", "rays_b200_m"):
            continue
        mod = m.group(1)
        path = None
        for d, _, files in os.walk(REF):
            for f in files:
                if f.lower() == mod.lower() + ".f90":
                    path = os.path.join(d, f)
        assert path, f"module {mod} not found in the reference"
        body = open(path, errors="replace").read().lower()
        for it in m.group(2).split(","):
            name = it.split("=>")[-1].strip().lower()
            assert re.search(r"\b%s\b" % re.escape(name), body), f"{mod}: no entity `{name}`"
            n += 1
    assert n > 100
