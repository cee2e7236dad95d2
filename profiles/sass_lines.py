"""Tuning aid (static): per-source-line instruction counts of one kernel of the built libc2rt.so.
usage: python profiles/sass_lines.py <kernel substring, e.g. 'ILi3ELi4E'> [opcode regex]
Needs cuobjdump + nvdisasm (CUDA toolkit); reads chess2rt_b200/libc2rt.so (or $C2RT_LIB_DIR/libc2rt.so)."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    want = sys.argv[1]
    op_re = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    lib = os.path.join(os.environ.get("C2RT_LIB_DIR", os.path.join(ROOT, "chess2rt_b200")), "libc2rt.so")
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL)
        cubin = max((os.path.join(td, f) for f in os.listdir(td)), key=os.path.getsize)
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    src = open(os.path.join(ROOT, "chess2rt_b200", "csrc", "render_kernel.cu")).read().splitlines()
    on, line, per_line, per_op, total = False, 0, collections.Counter(), collections.Counter(), 0
    for l in text.splitlines():
        if l.startswith(".text."):
            on = want in l
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File ".*", line (\d+)', l)
        if m:
            line = int(m.group(1))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
        if m:
            op = m.group(1)
            if op_re and not op_re.search(op):
                continue
            total += 1
            per_line[line] += 1
            per_op[op.split(".")[0]] += 1
    print("instructions:", total)
    print("top opcodes:", ", ".join(f"{k} {v}" for k, v in per_op.most_common(12)))
    for ln, n in per_line.most_common(60):
        print(f"{n:5d}  {ln:5d}  {src[ln - 1].strip()[:130] if 0 < ln <= len(src) else ''}")


if __name__ == "__main__":
    main()
