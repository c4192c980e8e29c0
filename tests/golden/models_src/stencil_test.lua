-- Test model written for this repository (not part of the reference).
-- Stencil / mask material overrides on top of painted and unpainted brushes.
local wedge = box(3, 0.4, 3):rotate_z(25)
local wedges = wedge:union(wedge:rotate_z(60)):union(wedge:rotate_z(120))
local body = sphere(3.2):paint("#f08020"):stencil(wedges, solid_material("#2040c0"))
local cap = cylinder(1.6, 3.6):mask(sphere(2.0):move(0, 0, 1.8), pbrbr_material("#30c060"))
model = body:diff(sphere(3.0):move(0, 0, 2.4)):union(cap:move(1.2, 0.4, 0)):blend_union(torus(4.2, 0.5):paint("#d0d0d0"), 0.25)
