# viscosity and density of water, SI units with mm (graphnics/flowmodels/parameters.py:1-2)
water_properties = {'nu': 0.697, 'rho': 1e-3, 'mu': 0.697 * 1e-3}
