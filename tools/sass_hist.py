#!/usr/bin/env python
"""usage: sass_hist.py <object or .so> <substring of the mangled kernel name> <out.txt>
Writes the SASS of one kernel (one instruction per line, encodings stripped) and prints its opcode histogram.
Runs on the CPU box: cuobjdump needs no GPU."""
import re,sys,subprocess
from collections import Counter
obj,pat=sys.argv[1],sys.argv[2]
out=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout.split('\n')
start=[i for i,l in enumerate(out) if 'Function :' in l and pat in l][0]
end=([i for i,l in enumerate(out) if 'Function :' in l and i>start]+[len(out)])[0]
body=[l for l in out[start:end] if re.search(r'/\*[0-9a-f]{4}\*/',l)]
ops=[re.sub(r'/\*[0-9a-f]+\*/','',l).split(';')[0].strip() for l in body]
open(sys.argv[3],'w').write('\n'.join(ops))
c=Counter()
for o in ops:
    t=o.split()
    if not t: continue
    m=t[1] if t[0].startswith('@') else t[0]
    c[m.split('.')[0]]+=1
print(len(ops)); print(c.most_common(22))
