"""Static issue-time estimate of a SASS range: decodes the scheduling control fields (stall count, yield, write / read
barrier, wait mask) that nvdisasm --print-instruction-encoding shows in the upper instruction word, and prints them next
to each instruction.  Development aid for the one-warp-per-scheduler entropy loops (k_seq, k_huf), whose time per step
is the sum of the stall counts on the path plus the scoreboard waits.

usage: sass_sched.py file.sass first_line last_line   (line numbers within the file, 1-based)
"""
import re
import sys


def parse(path, lo, hi):
    lines = open(path).read().split("\n")[lo - 1:hi]
    out = []
    i = 0
    while i < len(lines):
        m = re.search(r"/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            if m2:
                hi64 = int(m2.group(1), 16)
                ctrl = hi64 >> 41
                out.append(dict(addr=m.group(1), text=" ".join(m.group(2).split()), stall=ctrl & 15, yld=(ctrl >> 4) & 1,
                                wbar=(ctrl >> 5) & 7, rbar=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 63))
                i += 2
                continue
        if lines[i].strip().startswith(".L_"):
            out.append(dict(label=lines[i].strip()))
        i += 1
    return out


if __name__ == "__main__":
    ins = parse(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    tot = 0
    for x in ins:
        if "label" in x:
            print(x["label"])
            continue
        tot += x["stall"]
        wb = "" if x["wbar"] == 7 else "W%d" % x["wbar"]
        rb = "" if x["rbar"] == 7 else "R%d" % x["rbar"]
        wm = "".join(str(b) for b in range(6) if (x["wait"] >> b) & 1)
        print("%5s %4d s%-2d %1s %-3s %-3s wait[%-6s] %s" % (x["addr"], tot, x["stall"], "Y" if x["yld"] else "", wb, rb, wm, x["text"]))
    print("instructions", sum(1 for x in ins if "addr" in x), "stall sum", tot)
