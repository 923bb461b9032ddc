"""Find and load the UNMODIFIED reference modules of the moment-pooling path (benchmark comparator
and drop-in acceptance test only - the product package never imports this).

Lookup order: $EGM_REFERENCE, /root/reference (the build container), then
baseline/_ref (the git-ignored "install" that travels to the GPU box; `install()` below creates it).
The reference is pure Python without packaging metadata - `pip install /root/reference` stops with
"Neither 'setup.py' nor 'pyproject.toml' found" - so its install is a verbatim copy of `src/` plus a
sha256 manifest; nothing under baseline/_ref is edited, tracked by git or imported by the package.

`import src.models` itself fails in this image (timm / matplotlib are absent), so the
files are loaded by path into a synthetic package, with a caller-supplied stand-in for the timm
backbone module only (`cle_vit_backbone`), exactly as SURVEY.md 0.12 prescribes.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys
import types
from typing import Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALL_DIR = os.path.join(ROOT, "baseline", "_ref")
_PROBE = os.path.join("src", "models", "moment_head.py")


def find_reference_root() -> Optional[str]:
    for cand in (os.environ.get("EGM_REFERENCE"), "/root/reference", INSTALL_DIR):
        if cand and os.path.isfile(os.path.join(cand, _PROBE)):
            return cand
    return None


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def install(source: str = "/root/reference") -> Optional[str]:
    """Copy `<source>/src/**/*.py` verbatim to baseline/_ref/src and write MANIFEST.json (relative path
    -> sha256 of the source file). Returns the install dir, or None when the source is absent."""
    if not os.path.isfile(os.path.join(source, _PROBE)):
        return None
    manifest = {}
    for dirpath, _, files in os.walk(os.path.join(source, "src")):
        for name in sorted(files):
            if not name.endswith(".py"):
                continue
            src = os.path.join(dirpath, name)
            rel = os.path.relpath(src, source)
            dst = os.path.join(INSTALL_DIR, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[rel] = _sha(src)
    with open(os.path.join(INSTALL_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": source, "files": manifest}, f, indent=1, sort_keys=True)
    return INSTALL_DIR


def verify_install() -> bool:
    """True when every file of baseline/_ref still has the sha256 recorded at install time."""
    try:
        with open(os.path.join(INSTALL_DIR, "MANIFEST.json")) as f:
            files = json.load(f)["files"]
    except (OSError, ValueError, KeyError):
        return False
    return all(os.path.isfile(os.path.join(INSTALL_DIR, rel)) and _sha(os.path.join(INSTALL_DIR, rel)) == h
               for rel, h in files.items())


def _load(pkgname: str, root: str, rel: str, name: str):
    spec = importlib.util.spec_from_file_location(f"{pkgname}.{name}", os.path.join(root, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"{pkgname}.{name}"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_path_modules(root: Optional[str] = None, pkgname: str = "egm_reference_src"):
    """(gpf_kernel, moment_head, ops) of the reference, loaded by file path."""
    root = root or find_reference_root()
    if root is None:
        raise FileNotFoundError("reference sources not found (EGM_REFERENCE, /root/reference, baseline/_ref)")
    pkg = types.ModuleType(pkgname); pkg.__path__ = []
    models = types.ModuleType(pkgname + ".models"); models.__path__ = []
    utils = types.ModuleType(pkgname + ".utils"); utils.__path__ = []
    sys.modules.setdefault(pkgname, pkg)
    sys.modules.setdefault(pkgname + ".models", models)
    sys.modules.setdefault(pkgname + ".utils", utils)
    gk = _load(pkgname, root, "src/models/gpf_kernel.py", "models.gpf_kernel")
    mh = _load(pkgname, root, "src/models/moment_head.py", "models.moment_head")
    ops = _load(pkgname, root, "src/utils/ops.py", "utils.ops")
    return gk, mh, ops


def load_model_class(backbone_cls, *, root: Optional[str] = None, pkgname: str = "egm_reference_model",
                     native: bool = False):
    """The reference's `EGOMomentCLEViT` (ego_moment_clevit.py, loaded unchanged by file path) with
    `backbone_cls` standing in for `cle_vit_backbone.CLEViTDualStream` (timm is absent here).

    native=False: its `.gpf_kernel` / `.moment_head` are the reference's own files (the comparator).
    native=True : they are this repository's drop-in modules (`dropin.install_into`) - the reference's
                  model file and classifier head still run unchanged on top of them."""
    import importlib
    root = root or find_reference_root()
    if root is None:
        raise FileNotFoundError("reference sources not found (EGM_REFERENCE, /root/reference, baseline/_ref)")
    pkg = types.ModuleType(pkgname); pkg.__path__ = []
    models = types.ModuleType(pkgname + ".models"); models.__path__ = []
    sys.modules[pkgname] = pkg
    sys.modules[pkgname + ".models"] = models
    stub = types.ModuleType(pkgname + ".models.cle_vit_backbone")
    stub.CLEViTDualStream = backbone_cls
    sys.modules[pkgname + ".models.cle_vit_backbone"] = stub
    if native:
        importlib.import_module("ego-moment-cle-vit_b200").install_into(pkgname)
    else:
        _load(pkgname, root, "src/models/gpf_kernel.py", "models.gpf_kernel")
        _load(pkgname, root, "src/models/moment_head.py", "models.moment_head")
    _load(pkgname, root, "src/models/classifier_head.py", "models.classifier_head")
    mod = _load(pkgname, root, "src/models/ego_moment_clevit.py", "models.ego_moment_clevit")
    return mod.EGOMomentCLEViT


if __name__ == "__main__":
    print(install() or "reference source tree absent; nothing installed")
