"""Stand-in for torch-geometric 1.7.0 (only what the reference's hot-path modules touch)."""
__version__ = '1.7.0'
