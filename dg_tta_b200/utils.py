"""Process-global switch for the trainers' internal GIN augmentation.

Mirrors dg_tta/utils.py:5-18 of the reference (same function names, same environment variable,
same semantics — including the AttributeError when the variable was never set, utils.py:17-18).
"""
import os

_FLAG = "DG_TTA_INTERNAL_AUGMENTATION"


def enable_internal_augmentation():
    os.environ[_FLAG] = "true"


def disable_internal_augmentation():
    os.environ[_FLAG] = "false"


def check_internal_augmentation_disabled():
    assert os.environ.get(_FLAG).lower() != "true"


def get_internal_augmentation_enabled():
    return os.environ.get(_FLAG).lower() == "true"
