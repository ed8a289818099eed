"""PassiveTracer element type (opendrift/elements/passivetracer.py): LagrangianArray without extras."""
from .elements import LagrangianArray


class PassiveTracer(LagrangianArray):
    variables = LagrangianArray.add_variables([])
