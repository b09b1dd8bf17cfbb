"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Recipe that lets the LIVE Python reference travel to the GPU box. It copies, byte for byte and without
touching them, exactly the reference modules the hot path imports (the import closure of
``oracle/ref_harness.load_reference``: ``optimax_rogue/{game,logic,networking}`` + ``optimax_rogue_bots``
bot classes, about two dozen ``.py`` files) from ``/root/reference`` into ``oracle/_ref/``.

``oracle/_ref/`` is git-ignored (reference sources never enter the history) but not gpurun-ignored, so it
ships with the snapshot like the built ``.so`` files. On the box ``ref_harness`` finds it, which gives

  * ``bench.py --impl reference`` / ``cpu_baseline``: the reference's own ``Updater.update`` loop
    (``optimax_rogue/server/main.py:110-113``) timed on the box's host cores, ``kind: "reference"``;
  * ``tests/test_gpu_vs_reference.py``: CUDA against the live reference in ONE process.

    python -m oracle.make_ref            # run in the build container; __graft_entry__.build() calls it
"""
import json
import os
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(_HERE, '_ref')
SOURCE = '/root/reference'
MANIFEST = os.path.join(DEST, 'MANIFEST.json')


def import_closure(root):
    """Files under ``root`` that ``load_reference`` pulls in, found by importing it in a child interpreter."""
    import subprocess
    code = (
        'import json, os, sys\n'
        f'sys.path.insert(0, {os.path.dirname(_HERE)!r})\n'
        f'os.environ["ORX_REFERENCE_ROOT"] = {root!r}\n'
        'from oracle import ref_harness as rh\n'
        'rh.load_reference()\n'
        f'root = os.path.realpath({root!r}) + os.sep\n'
        'files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())\n'
        '                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(root)})\n'
        'print(json.dumps([f[len(root):] for f in files]))\n')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def make(source=SOURCE, dest=DEST, quiet=False):
    if not os.path.isdir(os.path.join(source, 'optimax_rogue', 'logic')):
        raise RuntimeError(f'no reference tree at {source}')
    files = import_closure(source)
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    for rel in files:
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(source, rel), dst)
    with open(os.path.join(dest, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': source, 'files': files,
                   'note': 'verbatim copies made by oracle/make_ref.py; git-ignored, shipped to the GPU box only'}, f, indent=1)
    if not quiet:
        print(f'{len(files)} reference modules -> {dest}')
    return files


def present(dest=DEST):
    return os.path.isdir(os.path.join(dest, 'optimax_rogue', 'logic'))


if __name__ == '__main__':
    make()
