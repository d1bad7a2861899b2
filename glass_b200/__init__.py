"""glass_b200 -- B200-native GLASS labeled message-passing hot path."""
__version__ = "0.1.0"
