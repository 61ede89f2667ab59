import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"])
    except Exception as e: print(f, "ERR", e)
