import re,sys,collections
lines=open(sys.argv[1]).read().splitlines()
ins=[]
for l in lines:
    m=re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);",l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
# backward branches
for a,t in ins:
    m=re.search(r"BRA(?:\.\S+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)",t)
    if m and int(m.group(1),16)<a: print("loop %#x -> %#x (%d instrs)"%(int(m.group(1),16),a,(a-int(m.group(1),16))//16))
if len(sys.argv)>3:
    lo,hi=int(sys.argv[2],16),int(sys.argv[3],16)
    c=collections.Counter()
    for a,t in ins:
        if lo<=a<=hi:
            t=re.sub(r"^@!?U?P\d\s+","",t)
            op=t.split()[0]
            op=op.split('.')[0] if not op.startswith("IMAD") else (".".join(op.split('.')[:2]) if any(x in op for x in ("WIDE","MOV","IADD","SHL","HI")) else "IMAD")
            c[op]+=1
    tot=sum(c.values())
    for k,v in c.most_common(): print("%6d %s"%(v,k))
    print("total",tot)
