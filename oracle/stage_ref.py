"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the reference's own CPU implementation of the coarse TSDF prior
(/root/reference/third_parties/fusion.py, numba `parallel=True`: the CPU path BASELINE.json's north_star names) into
oracle/_ref/ so that it travels to the GPU box, where /root/reference does not exist.

    python oracle/stage_ref.py          # build container only; __graft_entry__.build() calls stage()

The file is copied byte for byte (sha256 recorded next to it); oracle/_ref/ is git-ignored like every other built
artefact -- the reference source never enters the repository's history.  `load_fusion()` imports the staged file
with the one third-party module it needs and this image lacks (scikit-image, used only by get_mesh /
get_point_cloud, which nothing here calls) stubbed out.  Only tests/ and bench.py's cpu_baseline / --impl reference
legs may call this module; the product path (bnv_fusion_b200/) never does.
"""
import hashlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/third_parties/fusion.py"
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "third_parties_fusion.py")


def stage():
    """Copy the reference file into oracle/_ref/ (no-op when /root/reference is absent, e.g. on the GPU box)."""
    if not os.path.exists(REF_SRC):
        return os.path.exists(DST)
    os.makedirs(DST_DIR, exist_ok=True)
    shutil.copyfile(REF_SRC, DST)
    with open(DST, "rb") as f, open(DST + ".sha256", "w") as g:
        g.write(hashlib.sha256(f.read()).hexdigest() + "  third_parties/fusion.py (unmodified copy)\n")
    return True


def available():
    return os.path.exists(DST)


def load_fusion():
    """Import the staged reference module (CPU mode: PyCUDA is absent, so FUSION_GPU_MODE = 0)."""
    if not available():
        raise FileNotFoundError("oracle/_ref/third_parties_fusion.py is missing: run `python oracle/stage_ref.py` in the "
                                "build container (needs /root/reference)")
    if "skimage" not in sys.modules:
        try:
            import skimage.measure  # noqa: F401
        except ImportError:
            m = types.ModuleType("skimage.measure")
            sys.modules["skimage"] = types.ModuleType("skimage")
            sys.modules["skimage.measure"] = m
            sys.modules["skimage"].measure = m
    spec = importlib.util.spec_from_file_location("bnv_ref_third_parties_fusion", DST)
    mod = importlib.util.module_from_spec(spec)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):      # the module prints a PyCUDA warning at import
        spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("staged" if stage() else "reference checkout not found; nothing staged")
