"""DataSpec -- the only piece of heal_swin.data the hot path reads (data_spec.py:5-11)."""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple, Union


@dataclass
class DataSpec:
    dim_in: Union[int, Tuple[int, int]]  # number of HEALPix pixels (single int) for the HP model
    f_in: int
    f_out: int
    base_pix: Optional[int]
    class_names: List[str] = field(default_factory=list)
