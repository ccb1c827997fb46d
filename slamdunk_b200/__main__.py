from .slamdunk import run

run()
