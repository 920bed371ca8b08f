"""Put the reference's own implementation of the path beside the oracle, so that it can be timed on the GPU box's host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (zillow/salve) is pure Python; /root/reference exists in the build
container but not on the GPU box.  This recipe imports the unmodified reference through oracle/ref_import.py (I/O stubs only),
asks the interpreter which `salve.*` source files that import really loaded, and copies exactly those files -- unmodified --
into oracle/_ref/ (git-ignored: the copies never enter the history; the directory travels to the GPU box with the tree, like the
built .so files).  `bench.py --impl reference` and the `cpu_baseline` leg then run the reference's render_bev_pair
(salve/utils/bev_rendering_utils.py:417-480) under multiprocessing.Pool, as scripts/render_dataset_bev.py:111-113 does, and report
kind "reference"; without oracle/_ref they fall back to the bit-identical port (oracle/bev_oracle.py, kind "port").

    python scripts/make_oracle_ref.py        (called by __graft_entry__.build() when /root/reference is present)
"""

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
# also needed by oracle/preprocess_oracle.py's pin (scripts/make_golden_preprocess.py): the reference's own transform classes
EXTRA = ["salve/utils/transform.py", "salve/utils/normalization_utils.py", "LICENSE"]


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "salve")):
        print("make_oracle_ref: /root/reference not present, nothing to do")
        return 0
    os.environ["SALVE_REFERENCE_ROOT"] = SRC
    sys.path.insert(0, ROOT)
    from oracle import ref_import

    ref_import.load()
    files = sorted(
        os.path.relpath(m.__file__, SRC)
        for k, m in list(sys.modules.items())
        if (k == "salve" or k.startswith("salve.")) and getattr(m, "__file__", None) and m.__file__.startswith(SRC + os.sep)
    )
    files += [f for f in EXTRA if os.path.exists(os.path.join(SRC, f))]
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in files:
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), d)
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write("Unmodified copies from zillow/salve made by scripts/make_oracle_ref.py (git-ignored, measurement only):\n")
        f.write("\n".join(files) + "\n")
    print(f"make_oracle_ref: {len(files)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
