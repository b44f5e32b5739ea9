"""Parse include/wf_stgcn.h into {name: (ret, [arg types])} using ctypes-like type tags."""
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "wf_stgcn.h")


def _tag(ctype):
    t = ctype.strip()
    if "*" in t:
        return "char_p" if t.replace(" ", "") == "constchar*" else "p"
    t = t.replace("const ", "").strip()
    return {"int": "i", "long long": "ll", "float": "f", "size_t": "sz", "double": "d", "void": "void"}[t]


def parse_header(path=HEADER):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(wf_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argl = []
        if args.strip() not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                ty = re.sub(r"\b\w+$", "", a).strip() if not a.endswith("*") else a
                argl.append(_tag(ty))
        out[name] = (_tag(ret), argl)
    return out
