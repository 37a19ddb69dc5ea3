"""Compares two fingerprint files of scripts/gpu_hash.py (bit-level regression)."""
import json, sys
a, b = json.load(open(sys.argv[1])), json.load(open(sys.argv[2]))
bad = [k for k in a if not k.endswith(".seconds") and a[k] != b.get(k)]
for k in bad:
    print("DIFF", k, a[k], b.get(k))
print("fingerprints: %d compared, %d differ" % (len([k for k in a if not k.endswith('.seconds')]), len(bad)))
sys.exit(1 if bad else 0)
