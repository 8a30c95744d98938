import json, sys
rows = [json.loads(l) for l in sys.stdin if l.startswith("{")]
print("  ".join(f"{d['shape']}/M{d['M']}:{d['us']:.2f}" for d in rows))
