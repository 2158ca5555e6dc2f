// Planar articulated-body environments (Swimmer, Hopper): added in planar_envs step; see DESIGN.md.
#pragma once
