"""Drop-in mirror of the reference's ``lib/modeling`` package (same builders, same call signatures)."""
from .loss import SetCriterion, build_loss
from .matcher import HungarianMatcher, PerFrameMatcher, build_matcher
from .model import SketchLocalizationModel, build_model
from .postprocess import postprocess
from .svanet import SVANet, build_svanet

__all__ = ["SVANet", "build_svanet", "SetCriterion", "build_loss", "PerFrameMatcher", "HungarianMatcher",
           "build_matcher", "SketchLocalizationModel", "build_model", "postprocess"]
