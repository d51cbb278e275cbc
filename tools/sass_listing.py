"""SASS listing + opcode histogram of one kernel of libirsgmcmc.so (cuobjdump, no GPU needed).
Usage: python tools/sass_listing.py <kernel name regex> <out file> [lib]"""
import collections
import re
import subprocess
import sys


def main(pattern, out, lib='irsgmcmc_b200/libirsgmcmc.so'):
    text = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    blocks = re.split(r'\n\s*Function : ', text)
    hits = [b for b in blocks[1:] if re.search(pattern, b.split('\n', 1)[0])]
    if not hits:
        raise SystemExit(f'no function matches {pattern}')
    with open(out, 'w') as f:
        for b in hits:
            name, body = b.split('\n', 1)
            lines = [re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l) for l in body.split('\n')]
            ins = [re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', l) for l in lines]
            ins = [(m.group(1), m.group(2).strip()) for m in ins if m]
            hist = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', i).split()[0] for _, i in ins)
            f.write(f'# cuobjdump -sass {lib}, function {name.strip()}\n# {len(ins)} instructions (sm_100a). Opcode histogram:\n')
            for op, n in hist.most_common(28):
                f.write(f'#   {op:32s} {n}\n')
            for a, i in ins:
                f.write(f'{a} {i}\n')
            f.write('\n')
    print(out, sum(1 for _ in open(out)), 'lines')


if __name__ == '__main__':
    main(*sys.argv[1:])
