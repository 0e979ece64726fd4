"""Static proof that adding an opt-in kernel variant did not change the kernels that are measured and tested by default:
dump the SASS of every instantiation of a kernel template from two builds of libtfpnp_b200.so and compare them instruction by
instruction (encodings included), matching instantiations by their leading template arguments.
    python tools/sass_identity.py old.so new.so conv3x3_tc2 [suffix-of-new-default-args, e.g. Lb0E]
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
    a, b = load(old, kernel), load(new, kernel)
    ok = True
    for k in sorted(a):
        same = b.get(k + suffix) == a[k]
        ok &= same
        print(f"{kernel}<{k}>  {len(a[k])} lines  {'IDENTICAL' if same else 'DIFFERENT'}")
    for k in sorted(b):
        if not any(k == x + suffix for x in a):
            print(f"{kernel}<{k}>  {len(b[k])} lines  new instantiation")
    print("default instantiations byte-identical:", ok)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
