"""Tuning aid (dynamic): executed warp instructions and stall samples of one profiled launch per SOURCE LINE and per device
function, by joining the report's SASS rows (ncu --page source) with the line table of the same build (nvdisasm -g).
usage: python profiles/ncu_lines.py <report.ncu-rep> [top-N lines]        (libc2rt.so must be the build that was profiled)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "chess2rt_b200", "csrc", "render_kernel.cu")


def line_table(kernel_mangled_part):
    lib = os.path.join(os.environ.get("C2RT_LIB_DIR", os.path.join(ROOT, "chess2rt_b200")), "libc2rt.so")
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL)
        cubin = max((os.path.join(td, f) for f in os.listdir(td)), key=os.path.getsize)
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    on, line, out = False, 0, []
    for l in text.splitlines():
        if l.startswith(".text."):
            on = kernel_mangled_part in l
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File ".*", line (\d+)', l)
        if m:
            line = int(m.group(1))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?)\s*;", l)
        if m:
            out.append((line, m.group(1)))
    return out


def functions(src_lines):
    """(first line, name) of every function-like definition, by a rough match on the source."""
    starts = []
    for i, l in enumerate(src_lines, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__)[^;]*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m:
            starts.append((i, m.group(1)))
    return starts


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kname = rows[0][1]
    m = re.search(r"render_frame_kernel<\(int\)(\d+), \(int\)(\d+)>", kname)
    part = f"render_frame_kernelILi{m.group(1)}ELi{m.group(2)}E"
    hdr = rows[1]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows[2:] if len(r) > iex and r[isrc].strip()]
    table = line_table(part)
    if len(table) != len(body):
        print(f"warning: {len(table)} instructions in the build, {len(body)} in the report — not the same build?", file=sys.stderr)
    src = open(SRC).read().splitlines()
    fstarts = functions(src)

    def func_of(line):
        name = "?"
        for s, n in fstarts:
            if s <= line:
                name = n
            else:
                break
        return name

    per_line, per_fn, smp_line, smp_fn = (collections.Counter() for _ in range(4))
    movs = collections.Counter()
    tot = tots = 0
    for (line, sass), r in zip(table, body):
        n, s = int(r[iex]), int(r[ismp])
        tot += n
        tots += s
        per_line[line] += n
        smp_line[line] += s
        f = func_of(line)
        per_fn[f] += n
        smp_fn[f] += s
        op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
        if op.startswith(("MOV", "IMAD.MOV", "FSEL", "SEL", "UMOV", "CS2R")):
            movs[f] += n
    print(f"{kname}: {tot} warp instructions, {tots} stall samples")
    print("\nper function: instructions (share) | samples (share) | of which moves/selects")
    for f, n in per_fn.most_common(30):
        print(f"  {f:22s} {n:12d} {100.0 * n / tot:5.1f} % | {100.0 * smp_fn[f] / max(tots, 1):5.1f} % | {100.0 * movs[f] / max(n, 1):4.0f} %")
    print("\nper line: instructions (share) | samples share | line | source")
    for ln, n in per_line.most_common(top):
        text = src[ln - 1].strip()[:110] if 0 < ln <= len(src) else ""
        print(f"  {n:11d} {100.0 * n / tot:5.1f} % | {100.0 * smp_line[ln] / max(tots, 1):5.1f} % | {ln:5d} | {text}")


if __name__ == "__main__":
    main()
