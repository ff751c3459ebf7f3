from .sphere_harmonics import SHEncoder
