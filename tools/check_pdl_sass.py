#!/usr/bin/env python
"""Static check of the programmatic-dependent-launch discipline in the built library.

A kernel launched with programmatic stream serialisation may start while its predecessor is still running; everything
the predecessor produces must be read AFTER `griddepcontrol.wait` (SASS: ACQBULK).  nvcc is free to hoist loads through
`const __restrict__` pointers (LDG.E.*.CONSTANT) above the wait -- it did in csmri_rows_inv, round 2 -- so this lists,
per kernel, every read-only-path global load that precedes the first ACQBULK.  Loads of module-scope tables (plain LDG
of a `__device__` array) are not flagged: they are never written by a kernel.

    python tools/check_pdl_sass.py [tfpnp_b200/libtfpnp_b200.so]      exit status 1 if any kernel is flagged"""
import re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tfpnp_b200", "libtfpnp_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
bad = {}
n_wait = 0
for block in sass.split("Function : ")[1:]:
    name, _, body = block.partition("\n")
    ins = [l.split("*/", 1)[1].strip() for l in body.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    w = next((i for i, s in enumerate(ins) if s.startswith("ACQBULK")), None)
    if w is None:
        continue
    n_wait += 1
    early = [s.split(";")[0] for s in ins[:w] if re.match(r"(@!?U?P\d+\s+)?LDG\.[A-Z0-9.]*CONSTANT", s)]
    if early:
        bad[name.strip()] = early
def demangle(n):
    try: return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:110]
    except OSError: return n
print(f"{n_wait} kernels execute griddepcontrol.wait; {len(bad)} read through a const __restrict__ pointer before it")
for k, v in bad.items():
    print(" ", demangle(k))
    for s in v[:6]: print("      ", s)
sys.exit(1 if bad else 0)
