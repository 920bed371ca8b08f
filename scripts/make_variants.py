"""Build -D variants of libsalve_bev.so into scratch/ for scripts/variant_bench.py (developer tool).
    python scripts/make_variants.py tag1:DEF1=V1,DEF2=V2 tag2:DEF=V ...      ->  scratch/lib_<tag>.so"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salve_b200 import build as b

os.makedirs("scratch", exist_ok=True)
for spec in sys.argv[1:]:
    tag, _, defs = spec.partition(":")
    out = os.path.join("scratch", f"lib_{tag}.so")
    b.build(force=True, defines=[d for d in defs.split(",") if d], out=out)
    print(out)
