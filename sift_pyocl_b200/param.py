"""Default parameters of the SIFT algorithm.

Same names and values as the reference's ``par`` (sift-src/param.py:43-79); they are part of the
public API (``sift.par.PeakThresh`` ...).  The CUDA library hard-codes the same constants
(csrc/siftb_api.cu) -- as in the reference, where the kernels hard-code ``Scales`` (image.cl:355),
changing them here after import does not re-parameterise the device code.
"""


class Enum(dict):
    """dict whose keys are also attributes (reference param.py:43-50)."""

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)


par = Enum(OctaveMax=100000,
           DoubleImSize=0,
           order=3,
           InitSigma=1.6,
           BorderDist=5,
           Scales=3,
           PeakThresh=255.0 * 0.04 / 3.0,
           EdgeThresh=0.06,
           EdgeThresh1=0.08,
           OriBins=36,
           OriSigma=1.5,
           OriHistThresh=0.8,
           MaxIndexVal=0.2,
           MagFactor=3,
           IndexSigma=1.0,
           IgnoreGradSign=0,
           MatchRatio=0.73,
           MatchXradius=1000000.0,
           MatchYradius=1000000.0,
           noncorrectlylocalized=0)
