"""B200-native superpixel-align hot path (overlap CSR, pooling, prior-weighted k-means)."""
__version__ = '0.1.0'
