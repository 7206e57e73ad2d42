"""Counts the Blackwell-specific SASS mnemonics per object file of libgsn_b200.so (B200_PROFILING.md, "What proves a
Blackwell-native kernel") and writes profiles/<name>.  python scripts/sass_summary.py [out-name]"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
MNEMONICS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR',
             'UTCATOMSWS', 'SYNCS', 'HMMA', 'HGMMA', 'LDGSTS', 'REDUX', 'ATOMS', 'VOTE', 'SHFL', 'POPC', 'FLO']


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'r2_sass_summary.txt'
    lines = ['SASS mnemonic counts per object of gsn_b200/libgsn_b200.so (cuobjdump -sass, sm_100a; built by gsn_b200/build.py)', '']
    for obj in sorted(glob.glob(os.path.join(ROOT, 'gsn_b200', '_obj', '*.o'))):
        sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        cnt = collections.Counter()
        kernels = re.findall(r'Function : (\S+)', sass)
        for m in re.finditer(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', sass, re.M):
            op = m.group(1)
            for k in MNEMONICS:
                if op.startswith(k):
                    cnt[k] += 1
        total = len(re.findall(r'^\s+/\*[0-9a-f]+\*/\s+\S', sass, re.M))
        lines.append(f'{os.path.basename(obj)}: {len(kernels)} kernels, {total} instructions')
        lines.append('   ' + '  '.join(f'{k}={cnt[k]}' for k in MNEMONICS if cnt[k]))
    out = os.path.join(ROOT, 'profiles', name)
    with open(out, 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
