"""ORACLE - TEST / BASELINE INFRASTRUCTURE ONLY.  Stages the reference itself next to the oracle.

    python oracle/build_ref.py            # /root/reference/*.py  ->  oracle/_ref/   (git-ignored build output)

The reference is 8 Python scripts (no setup.py / pyproject: nothing to pip-install, nothing to compile), and
`/root/reference` does not exist on the GPU box.  `oracle/_ref/` is the same kind of artefact as a compiled `.so`:
produced here by this recipe from the sources where they lie, listed in `.gitignore` (so it never enters the history)
but not in `.gpurunignore` (so it travels to the GPU box with the snapshot).  Consumers: `oracle/ref_loader.py` ->
`tests/test_gpu_reference_harness.py` (the reference's UNMODIFIED train.py / evaluate.py driving this repo's model),
`tests/test_oracle.py` (port vs reference) and `bench.py --impl reference` / `cpu_baseline` (`kind: "reference"`).
`__graft_entry__.build()` runs it whenever `/root/reference` is present."""
import os
import shutil

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["models.py", "train.py", "evaluate.py", "utils.py", "constants.py", "datasets.py"]


def build(verbose=True):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        print("oracle/_ref: staged", ", ".join(FILES), "from", SRC)
    return True


if __name__ == "__main__":
    build()
