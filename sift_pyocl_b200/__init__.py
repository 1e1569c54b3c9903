"""sift_pyocl_b200 -- B200-native drop-in for the keypoint path of sift_pyocl.

Same public names as the reference package (sift-src/__init__.py:29-33):
``SiftPlan``, ``MatchPlan``, ``LinearAlign``, ``par`` and ``version``.
"""
version = "0.4-b200.1"

from .param import par  # noqa: E402
from .plan import SiftPlan  # noqa: E402
from .match import MatchPlan  # noqa: E402
from .alignment import LinearAlign  # noqa: E402

__all__ = ["SiftPlan", "MatchPlan", "LinearAlign", "par", "version"]
