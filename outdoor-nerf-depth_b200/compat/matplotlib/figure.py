from matplotlib import _Missing
Figure = _Missing("matplotlib.figure.Figure")
