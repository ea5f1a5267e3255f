"""python -m bnv_fusion_b200.compat.run_e2e model=fusion_pointnet_model dataset=... trainer.checkpoint=...

Runs the reference's UNMODIFIED src/run_e2e.py on the B200 hot path (see compat/__init__.py).  The
reference checkout is located through $BNV_REFERENCE_ROOT."""
import os
import runpy
import sys

from . import install


def main():
    root = os.environ.get("BNV_REFERENCE_ROOT")
    if not root or not os.path.exists(os.path.join(root, "src", "run_e2e.py")):
        raise SystemExit("set BNV_REFERENCE_ROOT to a checkout of likojack/bnv_fusion")
    install(root)
    sys.argv[0] = os.path.join(root, "src", "run_e2e.py")
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
