// Restatement of lightsfm's sfm.hpp (robotics-upo/lightsfm; un-vendored third-party dependency of
// the reference: /root/reference/package.xml:31, CMakeLists.txt:38-41; no pinned version).
// TEST INFRASTRUCTURE ONLY: it exists so that /root/reference/src/{sfw_planner,costmap_model,
// trajectory}.cpp compile unmodified into oracle/_ref.  The formulas are the published
// Moussaid/Helbing social-force model as implemented upstream:
// desired force, averaged obstacle force over the agent's obstacle points, anisotropic pair
// social force, optional group forces, Euler position update with speed cap and goal pop.
// Parity at this boundary is UNPINNED: the reference holds no tests or vectors for it.
// Call sites in the reference: src/sfw_planner.cpp:592,594,697; types at :156,486,600-610,681-698.
#ifndef SFW_STUB_LIGHTSFM_SFM_HPP
#define SFW_STUB_LIGHTSFM_SFM_HPP

#include "angle.hpp"
#include "vector2d.hpp"
#include <cmath>
#include <list>
#include <unordered_map>
#include <vector>

namespace sfm {

struct Forces {
  utils::Vector2d desiredForce;
  utils::Vector2d obstacleForce;
  utils::Vector2d socialForce;
  utils::Vector2d groupGazeForce;
  utils::Vector2d groupCoherenceForce;
  utils::Vector2d groupRepulsionForce;
  utils::Vector2d groupForce;
  utils::Vector2d globalForce;
  utils::Vector2d robotSocialForce;
};

struct Parameters {
  Parameters()
      : forceFactorDesired(2.0), forceFactorObstacle(10.0),
        forceSigmaObstacle(0.2), forceFactorSocial(2.1),
        forceFactorGroupGaze(3.0), forceFactorGroupCoherence(2.0),
        forceFactorGroupRepulsion(1.0), lambda(2.0), gamma(0.35), n(2.0),
        nPrime(3.0), relaxationTime(0.5) {}
  double forceFactorDesired;
  double forceFactorObstacle;
  double forceSigmaObstacle;
  double forceFactorSocial;
  double forceFactorGroupGaze;
  double forceFactorGroupCoherence;
  double forceFactorGroupRepulsion;
  double lambda;
  double gamma;
  double n;
  double nPrime;
  double relaxationTime;
};

struct Goal {
  utils::Vector2d center;
  double radius;
};

struct Agent {
  Agent()
      : desiredVelocity(0.6), radius(0.35), cyclicGoals(false),
        teleoperated(false), antimove(false), linearVelocity(0.0),
        angularVelocity(0.0), groupId(-1) {}
  Agent(double linearVelocity, double angularVelocity)
      : desiredVelocity(0.6), radius(0.35), cyclicGoals(false),
        teleoperated(true), antimove(false), linearVelocity(linearVelocity),
        angularVelocity(angularVelocity), groupId(-1) {}
  Agent(const utils::Vector2d &position, const utils::Angle &yaw,
        double linearVelocity, double angularVelocity)
      : position(position), yaw(yaw), desiredVelocity(0.6), radius(0.35),
        cyclicGoals(false), teleoperated(true), antimove(false),
        linearVelocity(linearVelocity), angularVelocity(angularVelocity),
        groupId(-1) {}

  void move(double dt) {
    // teleoperated unicycle midpoint step
    double imd = linearVelocity * dt;
    utils::Vector2d inc(
        imd * std::cos(yaw.toRadian() + angularVelocity * dt * 0.5),
        imd * std::sin(yaw.toRadian() + angularVelocity * dt * 0.5));
    yaw += utils::Angle::fromRadian(angularVelocity * dt);
    position += inc;
    velocity.set(linearVelocity * yaw.cos(), linearVelocity * yaw.sin());
  }

  int id; // deliberately uninitialised upstream; harness sets it
  utils::Vector2d position;
  utils::Vector2d velocity;
  utils::Angle yaw;
  utils::Vector2d movement;
  double desiredVelocity;
  double radius;
  std::list<Goal> goals;
  bool cyclicGoals;
  bool teleoperated;
  bool antimove;
  double linearVelocity;
  double angularVelocity;
  int groupId;
  Forces forces;
  Parameters params;
  std::vector<utils::Vector2d> obstacles1;
  std::vector<utils::Vector2d> obstacles2;
};

struct Group {
  utils::Vector2d center;
  std::vector<unsigned> agents;
};

class Map; // the reference never passes a map

class SocialForceModel {
public:
  SocialForceModel(SocialForceModel const &) = delete;
  void operator=(SocialForceModel const &) = delete;
  ~SocialForceModel() {}

  static SocialForceModel &getInstance() {
    static SocialForceModel singleton;
    return singleton;
  }

#define SFM SocialForceModel::getInstance()

  std::vector<Agent> &computeForces(std::vector<Agent> &agents,
                                    Map *map = nullptr) const {
    std::unordered_map<int, Group> groups;
    for (unsigned i = 0; i < agents.size(); i++) {
      if (agents[i].groupId < 0)
        continue;
      groups[agents[i].groupId].agents.push_back(i);
      groups[agents[i].groupId].center += agents[i].position;
    }
    for (auto it = groups.begin(); it != groups.end(); ++it)
      it->second.center /= (double)(it->second.agents.size());

    for (unsigned i = 0; i < agents.size(); i++) {
      utils::Vector2d desiredDirection = computeDesiredForce(agents[i]);
      computeObstacleForce(agents[i], map);
      computeSocialForce(i, agents);
      computeGroupForce(i, desiredDirection, agents, groups);
      agents[i].forces.globalForce =
          agents[i].forces.desiredForce + agents[i].forces.socialForce +
          agents[i].forces.obstacleForce + agents[i].forces.groupForce;
    }
    return agents;
  }

  void computeForces(Agent &me, std::vector<Agent> &agents,
                     Map *map = nullptr) {
    Group mygroup;
    if (me.groupId != -1) {
      mygroup.agents.push_back(me.id);
      mygroup.center = me.position;
      for (unsigned i = 0; i < agents.size(); i++) {
        if (agents[i].id == me.id)
          continue;
        if (agents[i].groupId == me.groupId) {
          mygroup.agents.push_back(i);
          mygroup.center += agents[i].position;
        }
      }
      mygroup.center /= (double)mygroup.agents.size();
    }
    utils::Vector2d desiredDirection = computeDesiredForce(me);
    computeObstacleForce(me, map);
    computeSocialForce(me, agents);
    computeGroupForce(me, desiredDirection, agents, mygroup);
    me.forces.globalForce = me.forces.desiredForce + me.forces.socialForce +
                            me.forces.obstacleForce + me.forces.groupForce;
  }

  std::vector<Agent> &updatePosition(std::vector<Agent> &agents,
                                     double dt) const {
    for (unsigned i = 0; i < agents.size(); i++)
      updatePosition(agents[i], dt);
    return agents;
  }

  void updatePosition(Agent &agent, double dt) const {
    utils::Vector2d initPos = agent.position;
    utils::Angle initYaw = agent.yaw;
    if (agent.teleoperated) {
      agent.move(dt);
    } else {
      agent.velocity += agent.forces.globalForce * dt;
      if (agent.velocity.norm() > agent.desiredVelocity) {
        agent.velocity.normalize();
        agent.velocity *= agent.desiredVelocity;
      }
      agent.yaw = agent.velocity.angle();
      agent.position += agent.velocity * dt;
      agent.linearVelocity = agent.velocity.norm();
      agent.angularVelocity = (agent.yaw - initYaw).toRadian() / dt;
    }
    agent.movement = agent.position - initPos;
    if (!agent.goals.empty() &&
        (agent.goals.front().center - agent.position).norm() <=
            agent.goals.front().radius) {
      Goal g = agent.goals.front();
      agent.goals.pop_front();
      if (agent.cyclicGoals)
        agent.goals.push_back(g);
    }
  }

private:
  SocialForceModel() {}

  static double sq(double x) { return x * x; }

  utils::Vector2d computeDesiredForce(Agent &agent) const {
    utils::Vector2d desiredDirection;
    if (!agent.goals.empty() &&
        (agent.goals.front().center - agent.position).norm() >
            agent.goals.front().radius) {
      utils::Vector2d diff = agent.goals.front().center - agent.position;
      desiredDirection = diff.normalized();
      agent.forces.desiredForce =
          agent.params.forceFactorDesired *
          (desiredDirection * agent.desiredVelocity - agent.velocity) /
          agent.params.relaxationTime;
      agent.antimove = false;
    } else {
      agent.forces.desiredForce = -agent.velocity / agent.params.relaxationTime;
      agent.antimove = true;
    }
    return desiredDirection;
  }

  void computeObstacleForce(Agent &agent, Map *map) const {
    (void)map;
    if (agent.obstacles1.size() > 0 || agent.obstacles2.size() > 0) {
      agent.forces.obstacleForce.set(0, 0);
      for (unsigned i = 0; i < agent.obstacles1.size(); i++) {
        utils::Vector2d minDiff = agent.position - agent.obstacles1[i];
        double distance = minDiff.norm() - agent.radius;
        agent.forces.obstacleForce +=
            agent.params.forceFactorObstacle *
            std::exp(-distance / agent.params.forceSigmaObstacle) *
            minDiff.normalized();
      }
      for (unsigned i = 0; i < agent.obstacles2.size(); i++) {
        utils::Vector2d minDiff = agent.position - agent.obstacles2[i];
        double distance = minDiff.norm() - agent.radius;
        agent.forces.obstacleForce +=
            agent.params.forceFactorObstacle *
            std::exp(-distance / agent.params.forceSigmaObstacle) *
            minDiff.normalized();
      }
      agent.forces.obstacleForce /=
          (double)(agent.obstacles1.size() + agent.obstacles2.size());
    } else {
      agent.forces.obstacleForce.set(0, 0);
    }
  }

  static utils::Vector2d pairSocialForce(const Agent &agent,
                                         const Agent &other) {
    utils::Vector2d diff = other.position - agent.position;
    utils::Vector2d diffDirection = diff.normalized();
    utils::Vector2d velDiff = agent.velocity - other.velocity;
    utils::Vector2d interactionVector =
        agent.params.lambda * velDiff + diffDirection;
    double interactionLength = interactionVector.norm();
    utils::Vector2d interactionDirection =
        interactionVector / interactionLength;
    utils::Angle theta = interactionDirection.angleTo(diffDirection);
    double B = agent.params.gamma * interactionLength;
    double thetaRad = theta.toRadian();
    double forceVelocityAmount =
        -std::exp(-diff.norm() / B - sq(agent.params.nPrime * B * thetaRad));
    double forceAngleAmount =
        -theta.sign() *
        std::exp(-diff.norm() / B - sq(agent.params.n * B * thetaRad));
    utils::Vector2d forceVelocity = forceVelocityAmount * interactionDirection;
    utils::Vector2d forceAngle =
        forceAngleAmount * interactionDirection.leftNormalVector();
    return agent.params.forceFactorSocial * (forceVelocity + forceAngle);
  }

  void computeSocialForce(unsigned index, std::vector<Agent> &agents) const {
    Agent &agent = agents[index];
    agent.forces.socialForce.set(0, 0);
    for (unsigned i = 0; i < agents.size(); i++) {
      if (i == index)
        continue;
      utils::Vector2d f = pairSocialForce(agent, agents[i]);
      agent.forces.socialForce += f;
      if (i == 0)
        agent.forces.robotSocialForce = f;
    }
  }

  void computeSocialForce(Agent &agent, std::vector<Agent> &agents) const {
    agent.forces.socialForce.set(0, 0);
    for (unsigned i = 0; i < agents.size(); i++) {
      if (agents[i].id == agent.id)
        continue;
      agent.forces.socialForce += pairSocialForce(agent, agents[i]);
    }
  }

  void groupForceCommon(Agent &agent, unsigned selfIndex, bool selfIsIndex,
                        const utils::Vector2d &desiredDirection,
                        const std::vector<Agent> &agents,
                        const Group &group) const {
    // gaze
    utils::Vector2d com = group.center;
    com = (1.0 / (double)(group.agents.size() - 1)) *
          ((double)group.agents.size() * com - agent.position);
    utils::Vector2d relativeCom = com - agent.position;
    utils::Angle visionAngle = utils::Angle::fromDegree(90);
    double elementProduct = desiredDirection.dot(relativeCom);
    utils::Angle comAngle = utils::Angle::fromRadian(std::acos(
        elementProduct / (desiredDirection.norm() * relativeCom.norm())));
    if (comAngle > visionAngle) {
      double desiredDirectionSquared = desiredDirection.squaredNorm();
      double desiredDirectionDistance =
          elementProduct / desiredDirectionSquared;
      agent.forces.groupGazeForce = desiredDirectionDistance * desiredDirection;
      agent.forces.groupGazeForce *= agent.params.forceFactorGroupGaze;
    }
    // coherence
    com = group.center;
    relativeCom = com - agent.position;
    double distance = relativeCom.norm();
    double maxDistance = ((double)group.agents.size() - 1) / 2;
    agent.forces.groupCoherenceForce = relativeCom;
    double softenedFactor = agent.params.forceFactorGroupCoherence *
                            (std::tanh(distance - maxDistance) + 1) / 2;
    agent.forces.groupCoherenceForce *= softenedFactor;
    // repulsion
    for (unsigned i = 0; i < group.agents.size(); i++) {
      if (selfIsIndex ? (selfIndex == group.agents[i]) : (i == 0))
        continue;
      const Agent &o = agents.at(group.agents[i]);
      utils::Vector2d diff = agent.position - o.position;
      if (diff.norm() < agent.radius + o.radius)
        agent.forces.groupRepulsionForce += diff;
    }
    agent.forces.groupRepulsionForce *= agent.params.forceFactorGroupRepulsion;
    agent.forces.groupForce = agent.forces.groupGazeForce +
                              agent.forces.groupCoherenceForce +
                              agent.forces.groupRepulsionForce;
  }

  void computeGroupForce(unsigned index,
                         const utils::Vector2d &desiredDirection,
                         std::vector<Agent> &agents,
                         const std::unordered_map<int, Group> &groups) const {
    Agent &agent = agents[index];
    agent.forces.groupForce.set(0, 0);
    agent.forces.groupGazeForce.set(0, 0);
    agent.forces.groupCoherenceForce.set(0, 0);
    agent.forces.groupRepulsionForce.set(0, 0);
    if (groups.count(agent.groupId) == 0 ||
        groups.at(agent.groupId).agents.size() < 2)
      return;
    groupForceCommon(agent, index, true, desiredDirection, agents,
                     groups.at(agent.groupId));
  }

  void computeGroupForce(Agent &me, const utils::Vector2d &desiredDirection,
                         std::vector<Agent> &agents, Group &group) const {
    me.forces.groupForce.set(0, 0);
    me.forces.groupGazeForce.set(0, 0);
    me.forces.groupCoherenceForce.set(0, 0);
    me.forces.groupRepulsionForce.set(0, 0);
    if (group.agents.size() < 2)
      return;
    groupForceCommon(me, 0, false, desiredDirection, agents, group);
  }
};

} // namespace sfm
#endif
