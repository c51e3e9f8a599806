import sys, json
NS=6; ROWS=36
class Budget(Exception): pass
def search(sy,sz,budget=300000):
    off=[((r//NS)*sz+(r%NS)*sy)%16 for r in range(ROWS)]
    sizes=[10,10,10,6]
    used=[False]*ROWS
    result=[]
    nodes=[0]
    def build_pass(k, cur, hw0, hw1):
        nodes[0]+=1
        if nodes[0]>budget: raise Budget()
        n=len(cur)
        if n==sizes[k]:
            result.append(list(cur))
            if k==3 or build_pass(k+1, [], frozenset(), frozenset()): return True
            result.pop(); return False
        slot=n
        seen=set()
        for row in range(ROWS):
            if used[row] or off[row] in seen: continue
            seen.add(off[row])          # rows with equal offsets are interchangeable at this slot
            a0=[]; a1=[]
            for xp in range(3):
                lane=slot*3+xp
                (a0 if lane<16 else a1).append((off[row]+xp)%16)
            if any(c in hw0 for c in a0) or any(c in hw1 for c in a1): continue
            used[row]=True; cur.append(row)
            if build_pass(k, cur, hw0|frozenset(a0), hw1|frozenset(a1)): return True
            cur.pop(); used[row]=False
        return False
    try:
        ok=build_pass(0, [], frozenset(), frozenset())
    except Budget:
        return "budget"
    return result if ok else None
found={}; unknown=[]
for sy in range(16):
    for sz in range(16):
        r=search(sy,sz)
        if r=="budget": unknown.append((sy,sz))
        elif r: found[(sy,sz)]=r
print(len(found),"classes with conflict-free rowmaps; undecided:",len(unknown))
print(sorted(found))
json.dump({"%d,%d"%k:v for k,v in found.items()}, open('/tmp/rowmaps.json','w'))
