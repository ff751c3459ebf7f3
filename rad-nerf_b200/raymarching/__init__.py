from .raymarching import *
