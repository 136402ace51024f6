"""Argument / return containers at the boundary and ``pairwise_iou`` (detectron2/structures/boxes.py:243-275).

The drop-ins are duck-typed: every function of this package accepts the reference's own ``Boxes`` / ``Instances``
(anything with ``.tensor``, respectively ``get_fields()`` / attribute access / ``image_size``), and a maintainer who
wants the results in the reference's own classes says so once::

    fsg.structures.use_containers(detectron2.structures.Boxes, detectron2.structures.Instances)

Without that call the results come back in the two small stand-ins below, which carry just what the return path
needs (a tensor with a length; named per-instance fields with a common length, indexing, ``.to``)."""
from typing import Any, Dict, Tuple

import torch

from . import ops


def as_tensor(boxes) -> torch.Tensor:
    """``Boxes``-like (has ``.tensor``) or a tensor -> the (n,4) tensor."""
    return boxes.tensor if hasattr(boxes, "tensor") else boxes


def cat_tensors(boxes_list) -> torch.Tensor:
    """``Boxes.cat`` (boxes.py:212-228) for any mix of Boxes-likes / tensors, single-element shortcut included."""
    ts = [as_tensor(b) for b in boxes_list]
    return ts[0] if len(ts) == 1 else torch.cat(ts, dim=0)


class Boxes:
    """(n,4) fp32 XYXY tensor with a length (boxes.py:72-110: fp32 cast, empty input -> (0,4))."""

    def __init__(self, tensor: torch.Tensor):
        device = tensor.device if isinstance(tensor, torch.Tensor) else torch.device("cpu")
        tensor = torch.as_tensor(tensor, dtype=torch.float32, device=device)
        if tensor.numel() == 0:
            tensor = tensor.reshape((0, 4))
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def to(self, device) -> "Boxes":
        return Boxes(self.tensor.to(device))

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2, "Indexing on Boxes with {} failed to return a matrix!".format(item)
        return Boxes(b)

    def __len__(self) -> int:
        return self.tensor.shape[0]

    @property
    def device(self):
        return self.tensor.device


class Instances:
    """Named per-instance fields of one image (structures/instances.py): what the drop-ins return."""

    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        fields = object.__getattribute__(self, "_fields")
        if name not in fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return fields[name]

    def set(self, name: str, value: Any) -> None:
        if len(self._fields):
            assert len(self) == len(value), "Adding a field of length {} to a Instances of length {}".format(
                len(value), len(self))
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def get(self, name: str) -> Any:
        return self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, device) -> "Instances":
        return type(self)(self._image_size, **{k: (v.to(device) if hasattr(v, "to") else v)
                                               for k, v in self._fields.items()})

    def __getitem__(self, item) -> "Instances":
        return type(self)(self._image_size, **{k: v[item] for k, v in self._fields.items()})

    def __len__(self) -> int:
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")


_BOXES, _INSTANCES = Boxes, Instances


def use_containers(boxes_cls=None, instances_cls=None):
    """Classes the drop-ins build their results with (default: the stand-ins above).  ``boxes_cls(tensor)`` and
    ``instances_cls(image_size, **fields)`` are the only constructor forms used."""
    global _BOXES, _INSTANCES
    _BOXES = boxes_cls if boxes_cls is not None else Boxes
    _INSTANCES = instances_cls if instances_cls is not None else Instances


def make_boxes(tensor):
    return _BOXES(tensor)


def make_instances(image_size, **fields):
    return _INSTANCES(image_size, **fields)


def pairwise_iou(boxes1, boxes2) -> torch.Tensor:
    """IoU of all N x M pairs (boxes.py:243-275), materialised; bit-exact with the reference's fp32 result.
    The fused training path (``ops.match_anchors``) never builds this matrix."""
    return ops.pairwise_iou(as_tensor(boxes1), as_tensor(boxes2))
