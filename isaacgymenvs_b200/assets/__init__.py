"""Compiled articulation models.

The robot descriptions (MJCF/URDF) live in the reference tree (`/root/reference/assets`), which
does not exist on the GPU box, so the importer's output is committed here as JSON
(`compiled/*.json`, regenerate with `python -m isaacgymenvs_b200.assets.compile_assets`).  When a
reference-style asset root is available `load_asset_file` parses the XML directly instead.
"""
import os
from ..importer.model import Model, BuildOptions

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compiled")

# asset file (as named in the task YAMLs) -> compiled blob name
KNOWN = {
    "mjcf/nv_ant.xml": "ant",
    "mjcf/nv_humanoid.xml": "humanoid",
    "urdf/cartpole.urdf": "cartpole",
    "urdf/anymal_c/urdf/anymal_minimal.urdf": "anymal",
    "mjcf/open_ai_assets/hand/shadow_hand.xml": "shadow_hand",
    "urdf/objects/cube_multicolor.urdf": "cube",
}


def load_compiled(name) -> Model:
    with open(os.path.join(_DIR, name + ".json")) as f:
        return Model.from_json(f.read())


def load_asset_file(asset_root, asset_file, opts: BuildOptions = None) -> Model:
    """gym.load_asset(): parse from the XML when it exists, else fall back to the committed blob
    compiled from the same file with the options the reference task passes."""
    path = os.path.join(asset_root, asset_file)
    if os.path.exists(path):
        if path.endswith(".urdf"):
            from ..importer.urdf import load_urdf
            return load_urdf(path, opts)
        from ..importer.mjcf import load_mjcf
        return load_mjcf(path, opts)
    norm = os.path.normpath(path).replace("\\", "/")
    for key, blob in KNOWN.items():
        if norm.endswith(key):
            return load_compiled(blob)
    raise FileNotFoundError(f"asset {path} not found and no compiled model is registered for it")
