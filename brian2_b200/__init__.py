"""brian2_b200 -- a B200-native (sm_100a) simulation device for Brian2.

``import brian2_b200`` registers the device, after which ``set_device('b200')`` is a drop-in
beside ``runtime`` and ``cpp_standalone`` (registration pattern:
brian2/devices/cpp_standalone/device.py:2036-2037).
"""
from ._brian2_path import ensure_brian2_importable

ensure_brian2_importable()

from .device import B200Device, b200_device  # noqa: E402,F401
from .codeobject import B200CodeObject, B200HostCodeObject  # noqa: E402,F401
from .cuda_generator import CUDACodeGenerator  # noqa: E402,F401

__all__ = ["B200Device", "b200_device", "B200CodeObject", "B200HostCodeObject", "CUDACodeGenerator"]
