"""``listdir`` helper the reference imports from ``op.utils_train`` (dataset.py:7)."""
import os

_IMG_EXT = (".png", ".jpg", ".jpeg", ".bmp", ".webp")


def listdir(path):
    """All image files under ``path`` (recursive), sorted."""
    found = []
    for root, _, files in os.walk(path):
        for f in files:
            if f.lower().endswith(_IMG_EXT):
                found.append(os.path.join(root, f))
    return sorted(found)
