"""Which kernels of two builds of the library differ in their SASS?  Used to prove that an opt-in
experiment (a new kernel behind a switch) leaves every kernel of the GPU-tested default path
bit-identical:
    python tools/sass_diff.py other/libflowmap_b200.so [flowmap_b200/csrc/libflowmap_b200.so]
Runs in the build container (cuobjdump only, no GPU)."""
import hashlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def kernels(so):
    out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    table, name = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            # anonymous-namespace hashes depend on the translation unit's path: strip them
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu_[0-9a-f]+", r"\1", line.split(":", 1)[1].strip())
            table[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            table[name].append(re.sub(r"/\*.*?\*/", "", line).strip())
    return {k: (len(v), hashlib.md5("\n".join(v).encode()).hexdigest()) for k, v in table.items()}


def main():
    a = kernels(sys.argv[1])
    b = kernels(sys.argv[2] if len(sys.argv) > 2 else ROOT / "flowmap_b200" / "csrc" / "libflowmap_b200.so")
    same = [k for k in b if k in a and a[k] == b[k]]
    gone = {a[k]: k for k in a if k not in b}  # (length, hash) -> old name
    renamed = 0
    for k in sorted(b):
        if k not in a:
            if b[k] in gone:  # same SASS under another (mangled) name, e.g. a new template parameter
                print(f"renamed  {b[k][0]:5d}  {gone.pop(b[k])[:60]} -> {k[:60]}  (SASS identical)")
                renamed += 1
            else:
                print(f"new      {b[k][0]:5d}  {k[:110]}")
        elif a[k] != b[k]:
            print(f"changed  {a[k][0]:5d} -> {b[k][0]:5d}  {k[:100]}")
    for old_name in sorted(gone.values()):
        print(f"removed  {a[old_name][0]:5d}  {old_name[:110]}")
    print(f"{len(same) + renamed} kernels identical")


if __name__ == "__main__":
    main()
