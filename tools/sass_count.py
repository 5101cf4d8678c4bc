#!/usr/bin/env python3
"""Static SASS instruction counts per issue port for a cubin/.so/.o (offline, no GPU).

  python tools/sass_count.py file.cubin                    plain per-function counts
  python tools/sass_count.py file.cubin --poseidon KERNEL  dynamic counts of one Poseidon permutation:
      the rolled loops of poseidon::permute are weighted by their trip counts (full-round loop x8,
      paired partial-round loop x11, half loop x2, out-of-line sbox7_quad x24).

Ports (B200, measured in profiles/): "A" = ALU + FP64 (shared issue port, 2 clk per warp instruction),
"B" = FMA pipe (IMAD 2 clk, IMAD.WIDE with a 64-bit addend ~5.2 clk, IMAD.HI ~4.3 clk).
"""
import collections
import re
import subprocess
import sys

OTHER = {'NOP', 'EXIT', 'BRA', 'LDG', 'STG', 'LDC', 'S2R', 'LDCU', 'UMOV', 'RET', 'CALL', 'BSSY', 'BSYNC', 'ULDC', 'S2UR', 'LDL',
         'STL', 'LDS', 'STS', 'BAR', 'WARPSYNC', 'DEPBAR', 'UIADD3', 'ULOP3', 'USHF', 'UISETP', 'USEL', 'UIMAD', 'ULEA', 'R2UR',
         'UPLOP3', 'CS2R', 'BREAK', 'UMOV32I', 'LDSM', 'UP2UR', 'UR2UP', 'ULEPC', 'LEPC', 'MEMBAR', 'ERRBAR', 'CCTL'}


def classify(ins):
    toks = ins.split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    base = op.split('.')[0]
    if base in OTHER:
        return 'other'
    if base in ('IMAD', 'FFMA', 'FMUL', 'FADD', 'HFMA2', 'IMUL', 'FFMA2', 'HADD2', 'HMUL2'):
        if 'WIDE' in op:
            return 'fma_wide' if ins.rstrip().endswith('RZ') else 'fma_wide_acc'
        if '.HI' in op:
            return 'fma_hi'
        return 'fma'
    if base in ('DFMA', 'DADD', 'DMUL', 'DSETP'):
        return 'fp64'
    return 'alu'


def parse(path):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    fns = collections.OrderedDict()
    fn = None
    for l in out.splitlines():
        m = re.match(r'\s+Function : (\S+)', l)
        if m:
            fn = m.group(1)
            fns[fn] = []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m and fn:
            fns[fn].append((int(m.group(1), 16), m.group(2).strip()))
    return fns


def summarize(c):
    a = c['alu'] + c['fp64']
    b_clk = 2 * c['fma'] + 2 * c['fma_wide'] + 5.2 * c['fma_wide_acc'] + 4.3 * c['fma_hi']
    tot = sum(v for k, v in c.items() if k != 'other')
    return (f"alu {c['alu']:.0f} fp64 {c['fp64']:.0f} | fma {c['fma']:.0f} wide {c['fma_wide']:.0f} wide_acc {c['fma_wide_acc']:.0f} "
            f"hi {c['fma_hi']:.0f} | other {c['other']:.0f} | total {tot:.0f} | portA {2 * a:.0f} clk, portB {b_clk:.0f} clk, issue {tot + c['other']:.0f}")


def main():
    fns = parse(sys.argv[1])
    if len(sys.argv) > 3 and sys.argv[2] == '--poseidon':
        for fn, ins in fns.items():
            if sys.argv[3] not in fn:
                continue
            loops = []  # backward branches
            call_target = None
            for addr, s in ins:
                m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s+)?0x([0-9a-f]+)', s)
                if m and int(m.group(1), 16) < addr:
                    loops.append((int(m.group(1), 16), addr))
                m = re.search(r'CALL\.REL\.NOINC\s+0x([0-9a-f]+)', s)
                if m:
                    call_target = int(m.group(1), 16)
            loops.sort()
            # expected: outer half loop, full-round loop, partial loop (by start address)
            outer = [l for l in loops if any(o[0] > l[0] and o[1] < l[1] for o in loops)]
            inner = [l for l in loops if l not in outer and (call_target is None or l[0] < call_target)]
            w = {}
            if len(inner) >= 2:
                w[inner[0]] = 8
                w[inner[1]] = 11
            c = collections.Counter()
            for addr, s in ins:
                k = classify(s)
                weight = 1
                if call_target is not None and addr >= call_target:
                    weight = 24  # three sbox7_quad calls per full round, 8 full rounds
                else:
                    for rng, ww in w.items():
                        if rng[0] <= addr <= rng[1]:
                            weight = ww
                            break
                    else:
                        if outer and outer[0][0] <= addr <= outer[0][1]:
                            weight = 2
                c[k] += weight
            print(fn[:50], 'loops', [(hex(a), hex(b)) for a, b in loops], 'call', hex(call_target or 0))
            print('  per permutation:', summarize(c))
        return
    for fn, ins in fns.items():
        c = collections.Counter(classify(s) for _, s in ins)
        print(fn[:60], summarize(c))


if __name__ == '__main__':
    main()
