"""TEST INFRASTRUCTURE ONLY.  Copies the two `[Exp evaluation complete] {...}` lines of the reference's committed run log
results/2_main_table/final_with_insite.txt (:6 population SINDy, :2362 INSITE; seed 1, gamma 2, 1000/100/100) into
tests/golden/ref_logline_seed1.json: the wire format utils/results_utils.py:121-128 parses (run.py:120-121,
train_sindy.py:72-112).  Run in the build container (needs /root/reference)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference/results/2_main_table/final_with_insite.txt'


def main():
    lines = open(SRC).read().splitlines()
    out = {'source': 'results/2_main_table/final_with_insite.txt:6, :2362 and :2094'}
    # :2094 = population SINDy of the run of 2023-05-15 whose cached collection was seeded with 10 (a second known answer)
    for name, no in (('sindy', 6), ('insite', 2362), ('sindy_seed10', 2094)):
        line = lines[no - 1]
        assert '[Exp evaluation complete] {' in line, (no, line[:80])
        out[name] = line.split('[Exp evaluation complete] ')[1].strip()
    with open(os.path.join(ROOT, 'tests', 'golden', 'ref_logline_seed1.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print({k: v[:120] for k, v in out.items()})


if __name__ == '__main__':
    main()
