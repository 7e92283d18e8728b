DATA, AFFINE, TYPE, PATH, STEM = "data", "affine", "type", "path", "stem"
INTENSITY, LABEL, LOCATION = "intensity", "label", "location"
