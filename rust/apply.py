#!/usr/bin/env python3
"""Applies the GPU drop-in to a checkout of DXist/sparse_linear_assignment v0.1.5:

    python3 apply.py /path/to/sparse_linear_assignment
    SLA_B200_LIB_DIR=<dir that holds libsla_b200.so> cargo test

The patches under patches/ are line-addressed ed scripts (`diff -e` output: commands `Na`, `N,Mc`, `N,Md`, last hunk
first), so they carry only the NEW text and none of the crate's own source; this script interprets them (so that no
`ed` binary is needed), copies the new files (src/ffi.rs, src/device.rs, build.rs) and checks the crate version."""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def apply_ed(path, script):
    lines = open(path).read().split("\n")
    trailing = lines[-1] == ""
    if trailing:
        lines.pop()
    cmds = open(script).read().split("\n")
    i = 0
    while i < len(cmds):
        c = cmds[i]
        i += 1
        if not c:
            continue
        m = re.fullmatch(r"(\d+)(?:,(\d+))?([acd])", c)
        if not m:
            raise SystemExit(f"{script}: cannot parse ed command {c!r}")
        a, b, op = int(m.group(1)), int(m.group(2) or m.group(1)), m.group(3)
        text = []
        if op in "ac":
            while cmds[i] != ".":
                text.append(cmds[i])
                i += 1
            i += 1
        if op == "a":
            lines[a:a] = text
        elif op == "c":
            lines[a - 1:b] = text
        else:
            del lines[a - 1:b]
    open(path, "w").write("\n".join(lines) + ("\n" if trailing else ""))


def main():
    if len(sys.argv) != 2 or not os.path.isfile(os.path.join(sys.argv[1], "src", "ksparse.rs")):
        raise SystemExit("usage: python3 apply.py <checkout of sparse_linear_assignment v0.1.5>")
    crate = sys.argv[1]
    if 'version = "0.1.5"' not in open(os.path.join(crate, "Cargo.toml")).read():
        raise SystemExit("the ed scripts are line-addressed against v0.1.5")
    if os.path.exists(os.path.join(crate, "src", "device.rs")) or "mod device;" in open(os.path.join(crate, "src", "lib.rs")).read():
        raise SystemExit("this checkout is already patched (src/device.rs / `mod device;` present): the ed scripts are "
                         "line-addressed against the pristine v0.1.5 sources and must not be applied twice")
    for rel in ("src/ksparse.rs", "src/symmetric.rs", "src/lib.rs", "Cargo.toml"):
        apply_ed(os.path.join(crate, rel), os.path.join(HERE, "patches", os.path.basename(rel) + ".ed"))
    for rel in ("src/ffi.rs", "src/device.rs", "build.rs"):
        shutil.copy(os.path.join(HERE, rel), os.path.join(crate, rel))
    print("patched: src/ksparse.rs src/symmetric.rs src/lib.rs Cargo.toml; added: src/ffi.rs src/device.rs build.rs")


if __name__ == "__main__":
    main()
