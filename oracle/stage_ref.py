"""Stage the two reference hot-path files into the git-ignored oracle/_ref/ so the "reference files
verbatim on restated ops" CPU arm can also run on the GPU box (where /root/reference does not exist).
oracle/_ref/ is listed in .gitignore (never enters history) and is not gpurun-ignored.
"""
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))


def stage(src_root: str = "/root/reference") -> bool:
    files = ["model/point_net2.py", "model/project_to_2d.py"]
    if not all(os.path.isfile(os.path.join(src_root, f)) for f in files):
        return False
    for f in files:
        dst = os.path.join(_HERE, "_ref", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, f), dst)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not present; nothing staged")
