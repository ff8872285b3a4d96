"""B200-native Gaussian-splatting render path (internal package behind the drop-in modules)."""
