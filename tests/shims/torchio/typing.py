from typing import Callable, Optional, Sequence, Tuple, Union

TypeNumber = Union[int, float]
TypeTripletInt = Tuple[int, int, int]
TypeSpatialShape = Union[int, TypeTripletInt]
TypeKeys = Optional[Sequence[str]]
TypeCallable = Callable
