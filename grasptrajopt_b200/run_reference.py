"""Launcher that runs an UNMODIFIED script of the reference (e.g. ``examples/pybullet_gto_planning.py``) against this package:

    python -m grasptrajopt_b200.run_reference /path/to/GraspTrajOpt/examples/pybullet_gto_planning.py --robot panda --scene_type tabletop -d DATA

The reference's scripts import ``_init_paths`` (``examples/_init_paths.py:10-13``), which puts the checkout's root at the FRONT of
``sys.path`` -- unless that exact path string is already on it.  The launcher therefore seeds ``sys.path`` as

    [<examples dir>, <grasptrajopt_b200/compat>, ..., "<examples dir>/.."]

so that ``add_path`` is a no-op and ``optas`` / ``gto`` / ``mesh_to_sdf`` resolve to the CasADi-free packages in ``compat/``,
while ``data/configs`` and ``data/robots`` are still found in the checkout (``gto.utils.get_root_dir`` honours ``GTO_ROOT_DIR``).
Nothing in the reference tree is edited or copied.
"""
from __future__ import annotations

import os
import runpy
import sys


def seed_paths(script: str) -> str:
    """Prepare ``sys.path`` / the environment for ``script``; returns the checkout root."""
    from . import install_compat

    examples = os.path.dirname(os.path.abspath(script))
    root = os.path.abspath(os.path.join(examples, ".."))
    compat = install_compat()
    for p in (examples, compat):
        while p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, compat)
    sys.path.insert(0, examples)  # what `python script.py` would do: the script's own directory first (pybullet_api, utils, _init_paths)
    lib_path = os.path.join(examples, "..")  # the exact string examples/_init_paths.py:12 builds
    if lib_path not in sys.path:
        sys.path.append(lib_path)
    os.environ.setdefault("GTO_ROOT_DIR", root)
    return root


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m grasptrajopt_b200.run_reference <reference script> [script arguments ...]")
    script = os.path.abspath(argv[0])
    seed_paths(script)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
