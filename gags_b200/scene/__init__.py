"""Model-state surface of /root/reference/scene: GaussianModel and the camera types.
(The COLMAP/Blender `Scene` loader is data I/O — out of scope per SURVEY §2 rows 10-11.)"""
from .gaussian_model import GaussianModel  # noqa: F401
from .cameras import Camera, MiniCam  # noqa: F401
from .dataset_readers import read_sam_clip_feature  # noqa: F401
