"""qbot_b200 -- B200-native state-path backend for the qbot DSL."""
__version__ = "0.1.0"
