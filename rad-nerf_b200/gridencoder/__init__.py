from .grid import GridEncoder
