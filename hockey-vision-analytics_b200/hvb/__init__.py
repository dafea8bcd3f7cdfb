"""hvb — B200-native (sm_100a) hot path of hockey-vision-analytics behind the reference's call surface.

    from hvb import TeamClassifier, HybridTeamClassifier, Detector, B200InferenceSlicer, Detections

Importing the package never touches the GPU; the first object that needs it creates a libhvb
context and raises ``HvbError`` when there is no B200 or no built ``libhvb.so`` — there is no CPU
fallback (the CPU restatement lives in ``oracle/`` and is test infrastructure only).
"""
from ._ffi import HvbError, LIB_PATH  # noqa: F401
from .detections import Detections, crop_image  # noqa: F401


def __getattr__(name):
    # heavy modules (torch) are imported lazily
    if name in ("Context", "get_context", "LetterboxPlan"):
        from . import runtime
        return getattr(runtime, name)
    if name in ("Detector", "PLAYER_CLASS_ID", "GOALKEEPER_CLASS_ID"):
        from . import detect
        return getattr(detect, name)
    if name in ("B200InferenceSlicer", "OverlapFilter"):
        from . import slicer
        return getattr(slicer, name)
    if name == "HybridTeamClassifier":
        from .hybrid import HybridTeamClassifier
        return HybridTeamClassifier
    if name == "TeamClassifier":
        from .team import TeamClassifier
        return TeamClassifier
    if name == "SegmentationTeamClassifier":
        from .team_segmentation import SegmentationTeamClassifier
        return SegmentationTeamClassifier
    if name in ("VideoProcessor", "Config", "FrameResult"):
        from . import video
        return getattr(video, name)
    if name in ("ByteTrack", "MultiClipByteTrack", "DeviceByteTrack"):
        from . import tracker
        return getattr(tracker, name)
    raise AttributeError(name)
