import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[r for r in rows if r and r[0]=="ID"][0]
ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
skip=int(sys.argv[2]) if len(sys.argv)>2 else 0
for r in rows:
    if len(r)==len(hdr) and r[0].isdigit() and int(r[0])>=skip: print(r[0], r[ki][:70], r[hdr.index("Grid Size")], int(r[vi])/1000, "us")
