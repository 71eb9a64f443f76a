#!/usr/bin/env python
"""Static issue-cycle estimate of SASS regions from the control bits (stall field, bits [105:109) of
each 128-bit instruction; see /opt/skills/guides/B300_MICROARCH.md 'Instruction issue & scheduling').

    cuobjdump -sass file.cubin | python scripts/tools/sass_stalls.py [function-substring]

Prints, per function, every maximal straight-line run of >= 200 instructions (the unrolled MAS tile
bodies) with its instruction count, summed stall cycles and the scoreboard waits in it.
"""
import re
import sys

pat_fn = re.compile(r"Function : (\S+)")
pat_ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
pat_hi = re.compile(r"^\s*/\* (0x[0-9a-f]{16}) \*/")


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else ""
    fn = None
    ins = []          # (addr, text, stall, wait_mask)
    out = {}
    pending = None
    for line in sys.stdin:
        m = pat_fn.search(line)
        if m:
            fn = m.group(1)
            ins = out.setdefault(fn, [])
            continue
        m = pat_ins.search(line)
        if m:
            pending = (int(m.group(1), 16), m.group(2).strip())
            continue
        m = pat_hi.match(line)
        if m and pending:
            hi = int(m.group(1), 16)
            stall = (hi >> 41) & 0xF
            wait = (hi >> 52) & 0x3F
            ins.append((pending[0], pending[1], stall, wait))
            pending = None
    for fn, ins in out.items():
        if want not in fn:
            continue
        print(fn)
        run = []
        def flush():
            if len(run) >= 200:
                n = len(run)
                st = sum(max(i[2], 1) for i in run)
                waits = sum(1 for i in run if i[3])
                shfl = sum(1 for i in run if "SHFL" in i[1])
                setp = sum(1 for i in run if "FSETP" in i[1])
                print(f"  run @{run[0][0]:#x}: {n} instr, {st} stall-cycles ({st / n:.2f}/instr), {waits} with sb-wait, "
                      f"{shfl} SHFL, {setp} FSETP -> {st / max(shfl, 1):.1f} cyc/frame static")
        for i in ins:
            t = i[1]
            if re.search(r"\b(BRA|EXIT|BSYNC|BSSY|WARPSYNC|BAR|RET|CALL|VOTE|SYNCS)\b", t):
                flush()
                run = []
            else:
                run.append(i)
        flush()


if __name__ == "__main__":
    main()
