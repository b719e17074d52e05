"""``DeskewSettings`` -- the validated form of the ``deskew:`` config block.

Stands in for ``biahub.settings.DeskewSettings`` at
``shrimpy/preprocessing.py:138-140`` (constructed from the dict), ``:56``
(``model_dump`` filtered by callee signature) and ``:240-242`` (attribute
reads).  Keys are the ones of ``config/mda/mantis/dynatrack_demo.yaml:161-164``
plus the two that ``shrimpy/dynatrack/manager.py:297-299`` injects.
"""

from __future__ import annotations

from typing import Optional

from pydantic import BaseModel, ConfigDict, PositiveFloat, PositiveInt, field_validator, model_validator


class DeskewSettings(BaseModel):
    model_config = ConfigDict(extra="forbid")

    pixel_size_um: PositiveFloat
    ls_angle_deg: PositiveFloat
    px_to_scan_ratio: Optional[PositiveFloat] = None
    scan_step_um: Optional[PositiveFloat] = None
    keep_overhang: bool = False
    average_n_slices: PositiveInt = 3

    @model_validator(mode="before")
    @classmethod
    def _derive_ratio(cls, data):
        # ratio = pixel size / scan step, to three decimals (scripts/measure_psf.py:225)
        if isinstance(data, dict) and data.get("px_to_scan_ratio") is None:
            step, pixel = data.get("scan_step_um"), data.get("pixel_size_um")
            if step is None or pixel is None:
                raise ValueError("px_to_scan_ratio is not given, so pixel_size_um and scan_step_um are both required")
            data = dict(data)
            data["px_to_scan_ratio"] = round(float(pixel) / float(step), 3)
        return data

    @field_validator("ls_angle_deg")
    @classmethod
    def _angle_range(cls, v: float) -> float:
        if not 0 < v <= 45:
            raise ValueError("light-sheet angle must lie in (0, 45] degrees")
        return round(float(v), 2)

    @field_validator("px_to_scan_ratio")
    @classmethod
    def _ratio_rounding(cls, v: Optional[float]) -> Optional[float]:
        return None if v is None else round(float(v), 3)
