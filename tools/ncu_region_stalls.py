import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sections = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
for sec in sections:
    hdr = sec["hdr"]; ix = {h:i for i,h in enumerate(hdr)}; data = sec["data"]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("=====", sec["name"][:70], "instr", len(data), "samples", tot)
    for b in range(0, len(data), B):
        chunk = data[b:b+B]
        s = sum(int(r[ix["# Samples"]]) for r in chunk)
        ex = sum(int(r[ix["Instructions Executed"]]) for r in chunk)
        g = lambda k: sum(int(r[ix[k]]) for r in chunk)
        if s > tot*0.01:
            ops = {}
            for r in chunk:
                toks = r[ix["Source"]].split()
                op = toks[0] if toks else ''
                if op.startswith('@') and len(toks)>1: op = toks[1]
                op = op.split('.')[0]
                ops[op] = ops.get(op,0)+1
            top = sorted(ops.items(), key=lambda x:-x[1])[:5]
            print(f"[{b:6d}] smp {s:6d} ({100*s/tot:4.1f}%) exec {ex:9d} noinst {g('stall_no_inst'):5d} bar {g('stall_barrier'):5d} lsb {g('stall_long_sb'):5d} wait {g('stall_wait'):5d} ssb {g('stall_short_sb'):4d} sel {g('stall_selected'):5d} {top}")
