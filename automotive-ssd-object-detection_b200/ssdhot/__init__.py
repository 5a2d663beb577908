"""ssdhot -- B200-native SSD300 multibox post-backbone hot path."""
__version__ = "0.1.0"
