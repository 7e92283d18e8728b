from .evaluator import Evaluator
from .labeled_tensor import LabeledTensor
from .segmentation_evaluator import SegmentationEvaluator, confusion_counts
from .label_map_evaluator import LabelMapEvaluator
