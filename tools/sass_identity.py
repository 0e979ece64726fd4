"""Static proof that adding an opt-in kernel variant did not change the kernels that are measured and tested by default:
dump the SASS of every instantiation of a kernel template from two builds of libtfpnp_b200.so and compare them instruction by
instruction (encodings included), matching instantiations by their leading template arguments.
    python tools/sass_identity.py old.so new.so conv3x3_tc2 [new-suffix [old-suffix]]
new-suffix: mangled default value(s) of template arguments ADDED in the new build (e.g. Li0E for `int XF = 0`);
old-suffix: what it replaces at the end of the old names, if an argument was re-typed (e.g. Lb0E for `bool XF2 = false`).
"""
import collections
import re
import subprocess
import sys


def load(so, kernel):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    d, name = collections.defaultdict(list), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            k = re.search(kernel + r"I(.*?)EEvNS", m.group(1))
            name = k.group(1) if k else None
            continue
        if name is not None:
            d[name].append(re.sub(r"/\*[0-9a-f]{4}\*/", "", line))
    return d


def main():
    old, new, kernel = sys.argv[1:4]
    suffix = sys.argv[4] if len(sys.argv) > 4 else ""
    old_suffix = sys.argv[5] if len(sys.argv) > 5 else ""
    a, b = load(old, kernel), load(new, kernel)
    ok = True
    matched = set()
    for k in sorted(a):
        if old_suffix and not k.endswith(old_suffix):
            continue                      # an opt-in instantiation of the old build
        nk = (k[:-len(old_suffix)] if old_suffix else k) + suffix
        same = b.get(nk) == a[k]
        ok &= same
        matched.add(nk)
        print(f"{kernel}<{k}> -> <{nk}>  {len(a[k])} lines  {'IDENTICAL' if same else 'DIFFERENT'}")
    for k in sorted(b):
        if k not in matched:
            print(f"{kernel}<{k}>  {len(b[k])} lines  opt-in instantiation")
    print("default instantiations byte-identical:", ok)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
