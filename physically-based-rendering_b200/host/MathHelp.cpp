#include "MathHelp.h"


/** Degree to radians: deg * 3.14159265359 / 180 in double, rounded to float (MathHelp.cpp:9-11). */
cl_float MathHelp::degToRad( cl_float deg ) {
	return ( deg * MH_PI / 180.0f );
}


/** Radians to degree (MathHelp.cpp:238-240). */
cl_float MathHelp::radToDeg( cl_float rad ) {
	return ( rad * 180.0f / MH_PI );
}


/** Bounding box of a list of points (MathHelp.cpp:20-36). */
void MathHelp::getAABB( const vector<cl_float4>& vertices, glm::vec3* bbMin, glm::vec3* bbMax ) {
	glm::vec3 lo( vertices[0].x, vertices[0].y, vertices[0].z );
	glm::vec3 hi = lo;
	for( size_t i = 1; i < vertices.size(); i++ ) {
		const cl_float4& v = vertices[i];
		lo.x = ( lo.x < v.x ) ? lo.x : v.x;
		lo.y = ( lo.y < v.y ) ? lo.y : v.y;
		lo.z = ( lo.z < v.z ) ? lo.z : v.z;
		hi.x = ( hi.x > v.x ) ? hi.x : v.x;
		hi.y = ( hi.y > v.y ) ? hi.y : v.y;
		hi.z = ( hi.z > v.z ) ? hi.z : v.z;
	}
	*bbMin = lo;
	*bbMax = hi;
}


/** Bounding box of a list of boxes (MathHelp.cpp:46-65). */
void MathHelp::getAABB(
	const vector<glm::vec3>& bbMins, const vector<glm::vec3>& bbMaxs, glm::vec3* bbMin, glm::vec3* bbMax
) {
	if( bbMins.empty() ) {
		*bbMin = glm::vec3();
		*bbMax = glm::vec3();
		return;
	}
	glm::vec3 lo = bbMins[0], hi = bbMaxs[0];
	for( size_t i = 1; i < bbMins.size(); i++ ) {
		lo = glm::min( bbMins[i], lo );
		hi = glm::max( bbMaxs[i], hi );
	}
	*bbMin = lo;
	*bbMax = hi;
}


/** Surface area of a box (MathHelp.cpp:95-101). */
cl_float MathHelp::getSurfaceArea( const glm::vec3& bbMin, const glm::vec3& bbMax ) {
	const cl_float dx = fabsf( bbMax.x - bbMin.x ), dy = fabsf( bbMax.y - bbMin.y ), dz = fabsf( bbMax.z - bbMin.z );
	const cl_float xy = dx * dy;
	const cl_float zy = dz * dy;
	const cl_float xz = dx * dz;
	return 2.0f * ( xy + zy + xz );
}


/** Index of the longest axis (MathHelp.cpp:185-194). */
short MathHelp::longestAxis( glm::vec3 bbMin, glm::vec3 bbMax ) {
	glm::vec3 sides = bbMax - bbMin;
	if( sides[0] > sides[1] ) {
		return ( sides[0] > sides[2] ) ? 0 : 2;
	}
	return ( sides[1] > sides[2] ) ? 1 : 2;
}


/** Phong tessellate a barycentric point (MathHelp.cpp:211-224). */
glm::vec3 MathHelp::phongTessellate(
	const glm::vec3 p1, const glm::vec3 p2, const glm::vec3 p3,
	const glm::vec3 n1, const glm::vec3 n2, const glm::vec3 n3,
	const float alpha, const float u, const float v
) {
	const float w = 1.0f - u - v;
	const glm::vec3 pBary = p1 * u + p2 * v + p3 * w;
	const glm::vec3 pTessellated =
		u * MathHelp::projectOnPlane( pBary, p1, n1 ) +
		v * MathHelp::projectOnPlane( pBary, p2, n2 ) +
		w * MathHelp::projectOnPlane( pBary, p3, n3 );
	return ( 1.0f - alpha ) * pBary + alpha * pTessellated;
}


/** MathHelp.cpp:227-229 */
glm::vec3 MathHelp::projectOnPlane( glm::vec3 q, glm::vec3 p, glm::vec3 n ) {
	return q - glm::dot( q - p, n ) * n;
}


void MathHelp::triCalcAABB( Tri* tri, const vector<cl_float4>* vertices, const vector<cl_float4>* normals ) {
	MathHelp::triCalcAABB( tri, vertices, normals, Cfg::get().value<float>( Cfg::RENDER_PHONGTESS ) );
}


/**
 * AABB of a face, grown for Phong tessellation when render.phong_tessellation > 0 and the three
 * vertex normals differ (MathHelp.cpp:250-310).
 */
void MathHelp::triCalcAABB(
	Tri* tri, const vector<cl_float4>* vertices, const vector<cl_float4>* normals, float alpha
) {
	const cl_float4& a = ( *vertices )[tri->face.x];
	const cl_float4& b = ( *vertices )[tri->face.y];
	const cl_float4& c = ( *vertices )[tri->face.z];

	glm::vec3 lo( a.x, a.y, a.z ), hi( a.x, a.y, a.z );
	const cl_float4* rest[2] = { &b, &c };
	for( int i = 0; i < 2; i++ ) {
		const cl_float4& v = *rest[i];
		lo.x = ( lo.x < v.x ) ? lo.x : v.x;
		lo.y = ( lo.y < v.y ) ? lo.y : v.y;
		lo.z = ( lo.z < v.z ) ? lo.z : v.z;
		hi.x = ( hi.x > v.x ) ? hi.x : v.x;
		hi.y = ( hi.y > v.y ) ? hi.y : v.y;
		hi.z = ( hi.z > v.z ) ? hi.z : v.z;
	}
	tri->bbMin = lo;
	tri->bbMax = hi;

	if( alpha <= 0.0f ) {
		return;
	}

	const glm::vec3 p1( a.x, a.y, a.z ), p2( b.x, b.y, b.z ), p3( c.x, c.y, c.z );
	const glm::vec3 n1 = FLOAT4_TO_VEC3( ( *normals )[tri->normals.x] );
	const glm::vec3 n2 = FLOAT4_TO_VEC3( ( *normals )[tri->normals.y] );
	const glm::vec3 n3 = FLOAT4_TO_VEC3( ( *normals )[tri->normals.z] );

	const glm::vec3 test = ( n1 - n2 ) + ( n2 - n3 );
	if( fabsf( test.x ) <= 0.000001f && fabsf( test.y ) <= 0.000001f && fabsf( test.z ) <= 0.000001f ) {
		return;
	}

	float thickness;
	glm::vec3 sidedropMin, sidedropMax;
	MathHelp::triThicknessAndSidedrop( alpha, p1, p2, p3, n1, n2, n3, &thickness, &sidedropMin, &sidedropMax );

	const glm::vec3 ng = glm::normalize( glm::cross( p2 - p1, p3 - p1 ) );
	const glm::vec3 p1thick = p1 + thickness * ng;
	const glm::vec3 p2thick = p2 + thickness * ng;
	const glm::vec3 p3thick = p3 + thickness * ng;

	tri->bbMin = glm::min( glm::min( tri->bbMin, p1thick ), glm::min( p2thick, p3thick ) );
	tri->bbMax = glm::max( glm::max( tri->bbMax, p1thick ), glm::max( p2thick, p3thick ) );
	tri->bbMin = glm::min( tri->bbMin, sidedropMin );
	tri->bbMax = glm::max( tri->bbMax, sidedropMax );
}


/** Thickness and side drops of the tessellated face (MathHelp.cpp:324-378). */
void MathHelp::triThicknessAndSidedrop(
	const float alpha,
	const glm::vec3 p1, const glm::vec3 p2, const glm::vec3 p3,
	const glm::vec3 n1, const glm::vec3 n2, const glm::vec3 n3,
	float* thickness, glm::vec3* sidedropMin, glm::vec3* sidedropMax
) {
	const glm::vec3 e12 = p2 - p1;
	const glm::vec3 e13 = p3 - p1;
	const glm::vec3 e23 = p3 - p2;
	const glm::vec3 e31 = p1 - p3;
	const glm::vec3 c12 = alpha * ( glm::dot( n2, e12 ) * n2 - glm::dot( n1, e12 ) * n1 );
	const glm::vec3 c23 = alpha * ( glm::dot( n3, e23 ) * n3 - glm::dot( n2, e23 ) * n2 );
	const glm::vec3 c31 = alpha * ( glm::dot( n1, e31 ) * n1 - glm::dot( n3, e31 ) * n3 );
	const glm::vec3 ng = glm::normalize( glm::cross( e12, e13 ) );

	const float k_tmp = glm::dot( ng, c12 - c23 - c31 );
	const float k = 1.0f / ( 4.0f * glm::dot( ng, c23 ) * glm::dot( ng, c31 ) - k_tmp * k_tmp );

	float u = k * (
		2.0f * glm::dot( ng, c23 ) * glm::dot( ng, c31 + e31 ) +
		glm::dot( ng, c23 - e23 ) * glm::dot( ng, c12 - c23 - c31 )
	);
	float v = k * (
		2.0f * glm::dot( ng, c31 ) * glm::dot( ng, c23 - e23 ) +
		glm::dot( ng, c31 + e31 ) * glm::dot( ng, c12 - c23 - c31 )
	);

	u = ( u < 0.0f || u > 1.0f ) ? 0.0f : u;
	v = ( v < 0.0f || v > 1.0f ) ? 0.0f : v;

	const glm::vec3 pt = MathHelp::phongTessellate( p1, p2, p3, n1, n2, n3, alpha, u, v );
	*thickness = glm::dot( ng, pt - p1 );

	static const float UV[9][2] = {
		{ 0.0f, 0.5f }, { 0.5f, 0.0f }, { 0.5f, 0.5f }, { 0.25f, 0.75f }, { 0.75f, 0.25f },
		{ 0.25f, 0.0f }, { 0.75f, 0.0f }, { 0.0f, 0.25f }, { 0.0f, 0.75f }
	};
	for( int i = 0; i < 9; i++ ) {
		const glm::vec3 p = MathHelp::phongTessellate( p1, p2, p3, n1, n2, n3, alpha, UV[i][0], UV[i][1] );
		if( i == 0 ) {
			*sidedropMin = p;
			*sidedropMax = p;
		}
		else {
			*sidedropMin = glm::min( *sidedropMin, p );
			*sidedropMax = glm::max( *sidedropMax, p );
		}
	}
}
