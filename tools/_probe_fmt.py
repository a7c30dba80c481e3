import sys,json
for ln in sys.stdin:
    if ln[:2] in ("1 ","2 ","4 ","8 "):
        w,rest=ln.split(" ",1); rows=json.loads(rest)
        print(w,[ (r["rank"],round(r["ms"],2),round(r["stages"]["digits"],2),round(r["stages"]["sort"],2)) for r in rows])
    elif ln.startswith("world"): print(ln.strip())
