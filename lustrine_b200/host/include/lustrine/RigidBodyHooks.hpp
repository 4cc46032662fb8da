// RigidBodyHooks.hpp — the three places where the C wrapper reaches INTO the rigid-body engine instead of calling a
// Lustrine::Bullet::* function (reference src/LustrineWrapper.cpp:494-523: a ray test on the dynamics world, and
// setGravity / setCollisionFlags on a btRigidBody).  The wrapper calls these hooks; whoever provides the
// Lustrine::Bullet functions provides them too: host/HostBodies.cpp for the built-in stand-in, and — when the
// library is built with LUSTRINE_B200_BULLET_HEADER pointing at the reference's BulletPhysics.hpp so that the
// reference's own BulletPhysics.cpp and a real Bullet world are linked in — a ten-line adapter
// (tests/cpp/real_bullet_hooks.cpp is one).
#pragma once

#include "glm_compat.hpp"

namespace Lustrine {
namespace Bullet {

struct Simulation;
void hook_set_body_gravity(Simulation* simulation, int body, glm::vec3 gravity);  // src/LustrineWrapper.cpp:514-518
void hook_set_body_no_collision_response(Simulation* simulation, int body);       // :520-523
int hook_is_grounded(Simulation* simulation, int body);                           // :494-503

}  // namespace Bullet
}  // namespace Lustrine
