"""Stage the UNMODIFIED reference package under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).

The contract's `pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` cannot work here: the
reference is a poetry project (build-backend poetry.core.masonry.api, absent from the offline wheelhouse; it also pins
python <= 3.11.9 while this image runs 3.12) - recorded in DESIGN.md.  The package is pure Python, so the install is a
plain copy of the three modules the hot path lives in (nvit/__init__.py, model.py, kohonen.py); train.py is left out
because it cannot be imported here or on the GPU box (kornia, dynaconf, wandb are not installed).  Nothing is edited.

    python baseline/stage_ref.py            # no-op when /root/reference is absent (the GPU box uses the staged copy)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/nvit"
DST = os.path.join(HERE, "_ref", "nvit")
FILES = ("__init__.py", "model.py", "kohonen.py")


def stage() -> bool:
    if not os.path.isdir(SRC):
        return os.path.isfile(os.path.join(DST, "model.py"))
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return True


def import_reference():
    """`nvit.model` of the staged reference (raises if it was never staged)."""
    root = os.path.join(HERE, "_ref")
    if not os.path.isfile(os.path.join(DST, "model.py")):
        raise RuntimeError("baseline/_ref/nvit is not staged: run `python baseline/stage_ref.py` where /root/reference exists")
    if root not in sys.path:
        sys.path.insert(0, root)
    import nvit.model as ref_model
    return ref_model


if __name__ == "__main__":
    print("staged" if stage() else "reference not available")
