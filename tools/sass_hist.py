import subprocess, sys, re, collections
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(['cuobjdump','-sass',lib],capture_output=True,text=True).stdout
cur=None; funcs=collections.OrderedDict()
for line in out.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: cur=m.group(1); funcs[cur]=[]; continue
    m=re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(.*?);',line)
    if m and cur: funcs[cur].append(m.group(1))
for f,ins in funcs.items():
    if pat in f:
        ops=collections.Counter()
        for i in ins:
            t=i.split()
            op=t[1] if t[0].startswith('@') else t[0]
            ops[op.split('.')[0]]+=1
        print(f,len(ins)); print(' ',dict(ops.most_common(25)))
