// real_bullet_hooks.cpp — TEST INFRASTRUCTURE: the three RigidBodyHooks of the C wrapper for a host library built
// against the reference's BulletPhysics.hpp (a real Bullet world; see tests/cpp/Makefile).  They do what the
// reference's wrapper does at src/LustrineWrapper.cpp:494-523 — a short ray test below the body, setGravity and the
// no-contact-response flag on the btRigidBody.
#include LUSTRINE_B200_BULLET_HEADER
#include "lustrine/RigidBodyHooks.hpp"

namespace Lustrine {
namespace Bullet {

void hook_set_body_gravity(Simulation* s, int body, glm::vec3 g) { s->rigidbodies[body]->setGravity(btVector3(g.x, g.y, g.z)); }

void hook_set_body_no_collision_response(Simulation* s, int body) {
    s->rigidbodies[body]->setCollisionFlags(btCollisionObject::CF_NO_CONTACT_RESPONSE);
}

int hook_is_grounded(Simulation* s, int body) {
    const glm::vec3 p = get_body_position(s, body);
    const btVector3 from(p.x, p.y, p.z), to(p.x, p.y - 0.55f, p.z);
    btCollisionWorld::ClosestRayResultCallback hit(from, to);
    s->dynamicWorld->rayTest(from, to, hit);
    return (int)hit.hasHit();
}

}  // namespace Bullet
}  // namespace Lustrine
