"""``listdir`` helper the reference imports from ``op.utils_train`` (dataset.py:7; /root/reference/op/utils_train.py:8-26).

Reference contract: ``listdir(path, list_name)`` appends the full path of EVERY file below ``path`` (recursive, no
extension filter) to the caller's list, leaves that list sorted, and returns None — dataset.py:49/246/397/459 call it as
``listdir(self.root, img_names)``.  Calling it without a list returns a new sorted list (convenience, not in the reference).
"""
import os


def listdir(path, list_name=None):
    found = list_name if list_name is not None else []
    for root, dirs, files in os.walk(path):
        dirs.sort()
        found.extend(os.path.join(root, f) for f in files)
    found.sort()
    return None if list_name is not None else found
